#!/bin/bash
# A/B of the matcher's fold / no-fold variants: parity suite under each forced variant (unless "quick"), then kernel and step times.
if [ "${1:-}" != "quick" ]; then
  for f in 0 1; do
    SSRLCV_TEST_FOLD=$f python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  done
  python -m pytest tests -m gpu -x -q 2>&1 | tail -2
else
  SSRLCV_TEST_FOLD=0 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
fi
for c in C3 C3U C2 C1; do
  for f in 0 1; do
    w=$c; abs=40000
    if [ $c = C3U ]; then w=C3; abs=3e38; fi
    SSRLCV_BENCH_ABS=$abs SSRLCV_BENCH_TUNE=fold=$f python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$c fold=$f step ms', round(d['ms_per_step'],4), 'kernel ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'matches', d['matches_per_step'], d['clocks'])
"
  done
done
