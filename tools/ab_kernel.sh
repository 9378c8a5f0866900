#!/bin/bash
# A/B of the matcher's variants at C3: fold (five MMAs) / no-fold (four MMAs + window bounds) x bookkeeping regime
# (column masks / windows only).  Each cell: kernel + step time from bench.py; then the parity suite under each variant.
tag=${1:-ab}
mkdir -p gpurun_out
for f in 1 0; do for cm in 1 0; do
  SSRLCV_BENCH_TUNE="fold=$f,colmask=$cm" python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('fold=$f colmask=$cm', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'step_ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'matches', d['matches_per_step'], d['clocks'])
" | tee -a gpurun_out/${tag}_ab.txt
done; done
if [ "$2" == "tests" ]; then
for f in 1 0; do for cm in 1 0; do
  SSRLCV_TEST_FOLD=$f SSRLCV_TEST_COLMASK=$cm python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | sed "s/^/fold=$f colmask=$cm: /" | tee -a gpurun_out/${tag}_ab.txt
done; done
fi
