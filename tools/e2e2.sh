mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "match_host" 2>&1 | tail -3
for t in "" "chunk_q=16384,chunk_t=16384" "chunk_q=32768,chunk_t=32768"; do
  SSRLCV_BENCH_TUNE="$t" python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('tune [$t]', 'e2e ms', round(d['e2e']['ms_per_step'],4), 'step', round(d['ms_per_step'],4))
" | tee -a gpurun_out/r2j_e2e.txt
done
python bench.py --workload C2 --no-cpu-baseline 2>/dev/null | cut -c1-250
