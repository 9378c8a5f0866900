for rep in 1 2; do
for c in "-1 -1" "16384 16384" "32768 16384" "16384 32768" "-1 -1"; do
  set -- $c
  if [ "$1" == "-1" ]; then T=""; else T="chunk_q=$1,chunk_t=$2"; fi
  SSRLCV_BENCH_TUNE="$T" python bench.py --no-cpu-baseline --steps 20 2> /tmp/err.txt | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('chunks $1 $2', 'e2e ms', round(d['e2e']['ms_per_step'],4), 'step', round(d['ms_per_step'],4))
"
done; done
