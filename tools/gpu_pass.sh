mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/r2b_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -60 > gpurun_out/r2b_pytest_gpu.log; tail -5 gpurun_out/r2b_pytest_gpu.log
for c in C3 C2 C1; do
  timeout 300 python bench.py --workload $c --no-cpu-baseline > gpurun_out/r2b_bench_$c.json 2> gpurun_out/r2b_bench_$c.err; echo "bench $c rc=$?"; cut -c1-400 gpurun_out/r2b_bench_$c.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches_C3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches_C1.csv python bench.py --workload C1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_under_ncu1.log 2>&1
