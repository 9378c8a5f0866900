mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_launches_C4.csv python bench.py --workload C4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2h_under_ncu_c4.log 2>&1
timeout 600 python bench.py --workload C4 --steps 5 --no-cpu-baseline > gpurun_out/r2h_bench_C4.json 2> gpurun_out/r2h_bench_C4.err; cut -c1-200 gpurun_out/r2h_bench_C4.json
timeout 600 python bench.py --workload C3 --no-cpu-baseline > gpurun_out/r2h_bench_C3.json 2> gpurun_out/r2h_bench_C3.err; cut -c1-200 gpurun_out/r2h_bench_C3.json
