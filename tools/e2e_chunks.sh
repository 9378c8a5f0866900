#!/bin/bash
# e2e time of match_host at C3 for several cuts of the streamed upload ("query rows,target rows"; 0 = whole set)
for c in "0 0" "0 16384" "0 8192" "0 32768" "32768 32768" "32768 16384" "16384 16384" "8192 8192"; do
  set -- $c
  SSRLCV_BENCH_TUNE="chunk_q=$1,chunk_t=$2" python bench.py --no-cpu-baseline --steps 10 2> /tmp/err.txt | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('chunks $1 $2', 'e2e ms', round(d['e2e']['ms_per_step'],4))
"
done
