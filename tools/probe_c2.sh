#!/bin/bash
# kernel-role cycle counters for the shapes behind C2 (masked match, seed pass) -- diagnostics only
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${1}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${1}_pytest_gpu.log
{
python tools/kerneltiming.py 16384 16384 4e4 2 25
python tools/kerneltiming.py 16384 16384 4e4 0
python tools/kerneltiming.py 16384 16384 4e4 2 1e9
python tools/kerneltiming.py 16384 16384 3e38 2 25
python tools/kerneltiming.py 16384 8192 3e38 0
} > gpurun_out/${1}_probe_c2.txt 2>&1
for c in C3 C2 C1; do python bench.py --workload $c --no-cpu-baseline > gpurun_out/${1}_bench_$c.json 2> gpurun_out/${1}_bench_$c.err; done
python tools/benchsum.py gpurun_out/${1}_bench_C*.json
