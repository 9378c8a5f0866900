#!/bin/bash
# compute-sanitizer over the smoke call and a slice of the parity suite that drives every kernel of the chain (both matcher walks,
# masked / sorted searches, the streamed host path, the runner-up extension).  -> gpurun_out/<tag>_sanitizer.txt
tag=${1:-r2}
out=gpurun_out/${tag}_sanitizer.txt
mkdir -p gpurun_out
K="match_host_streamed_masked or forced_matcher_variants or degenerate or runner_up or ties_everywhere or exact_baseline_shapes"
{
echo "# compute-sanitizer on a B200 box ($tag)"
for tool in memcheck synccheck; do
  echo "compute-sanitizer --tool $tool python __graft_entry__.py --smoke"
  timeout 600 compute-sanitizer --tool $tool python __graft_entry__.py --smoke 2>&1 | grep -E "smoke ok|ERROR SUMMARY|Invalid|Error" | head -8
done
echo "SSRLCV_TEST_PAIR=1 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k \"$K\""
SSRLCV_TEST_PAIR=1 timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" | head -12
K2="cta_pair_walk or without_seeds or ties_everywhere"
echo "compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k \"$K2\"   (the CTA-pair walk, the cell sort / tile gather, the cull kernel, every mask-evaluation mode)"
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "$K2" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" | head -12
# racecheck per kernel: it does not model mbarrier hand-offs, so the matcher (whose roles synchronise through mbarriers only) is left out
for k in cell_sort_kernel gather_tiles_kernel cull_kernel; do
  echo "compute-sanitizer --tool racecheck --kernel-name kernel_substring=$k python -m pytest tests -m gpu -x -q -k without_seeds"
  timeout 600 compute-sanitizer --tool racecheck --kernel-name kernel_substring=$k python -m pytest tests -m gpu -x -q -k "without_seeds" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" | head -4
done
echo "compute-sanitizer --tool initcheck python __graft_entry__.py --smoke"
timeout 600 compute-sanitizer --tool initcheck python __graft_entry__.py --smoke 2>&1 | grep -E "smoke ok|ERROR SUMMARY|Uninitialized" | head -6
} > $out 2>&1
cat $out
