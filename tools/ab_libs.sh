#!/bin/bash
# same-box A/B of library builds: ab_libs/*.so (git-ignored), each run through the masked fixture timings
for l in ab_libs/*.so; do
  echo "== $l"
  SSRLCV_B200_LIB=$PWD/$l timeout 300 python tools/fixture_kt.py 2 1 2>&1 | grep "matcher kernel"
  SSRLCV_B200_LIB=$PWD/$l timeout 300 python tools/mask_walk_sweep.py 2>&1 | tail -2
done
for l in ab_libs/*.so; do
  echo "== $l (second pass)"
  SSRLCV_B200_LIB=$PWD/$l timeout 300 python tools/fixture_kt.py 2 1 2>&1 | grep "matcher kernel"
done
