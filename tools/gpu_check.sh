#!/bin/bash
# One GPU-box pass: parity tests, headline benches (ours + reference arm), launch list and one full ncu capture of the
# matcher kernel.  Run from the repo root under gpurun; everything lands in gpurun_out/<tag>_*.
#   tools/gpu_check.sh <tag> [tests] [bench] [ncu] [prof]
set -u
tag=${1:-run}; shift || true
what=${*:-tests bench ncu prof}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/${tag}_smi.txt 2>&1
for w in $what; do
  case $w in
    tests)
      python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log ;;
    bench)
      for c in C3 C2 C1; do
        python bench.py --workload $c > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; echo "bench $c rc=$?"
      done
      python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_C3_reference.json 2> gpurun_out/${tag}_bench_C3_reference.err
      python bench.py --workload C4 --steps 5 > gpurun_out/${tag}_bench_C4.json 2> gpurun_out/${tag}_bench_C4.err
      python bench.py --workload C5 --steps 3 > gpurun_out/${tag}_bench_C5.json 2> gpurun_out/${tag}_bench_C5.err ;;
    prof)
      python tools/e2eprof.py C3 > gpurun_out/${tag}_e2eprof_C3.txt 2>&1
      python tools/e2eprof.py C1 > gpurun_out/${tag}_e2eprof_C1.txt 2>&1
      python tools/kerneltiming.py C3 > gpurun_out/${tag}_kerneltiming_C3.txt 2>&1 ;;
    ncu)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_C3.csv \
          python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
      for c in C1 C2; do
        ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$c.csv \
            python bench.py --workload $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu_$c.log 2>&1
      done
      ncu --set full --clock-control none --import-source on -k regex:smax_ -s 3 -c 1 -f -o gpurun_out/${tag}_smax_C3 \
          python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1 ;;
  esac
done
ls -la gpurun_out | tail -30
