#!/bin/bash
# The evidence pass of a round: smoke, parity suite, every bench line (ours + reference arm), launch lists, one full ncu capture of
# the matcher, the per-role cycle breakdown and the MMA-only microbenchmark.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/${tag}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1; tail -1 gpurun_out/${tag}_smoke.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest_gpu.log; tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench C3 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_c3_reference.json 2> gpurun_out/${tag}_bench_c3_reference.err; echo "reference rc=$?"
for c in C2 C1 C4 C5; do
  timeout 600 python bench.py --workload $c --no-cpu-baseline > gpurun_out/${tag}_bench_$(echo $c | tr A-Z a-z).json 2> gpurun_out/${tag}_bench_$c.err; echo "bench $c rc=$?"
done
for c in C3 C2 C1; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/${tag}_launches_$(echo $c | tr A-Z a-z).csv python bench.py --workload $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_under_ncu_$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smax_ -s 3 -c 1 -f -o gpurun_out/${tag}_smax_c3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out/${tag}_smax_c3.ncu-rep
python tools/kerneltiming.py C3 > gpurun_out/${tag}_kerneltiming_c3.txt 2>&1
[ -x tools/microbench ] && ./tools/microbench > gpurun_out/${tag}_microbench.txt 2>&1
python tools/e2eprof.py C3 > gpurun_out/${tag}_e2eprof_c3.txt 2>&1
python - <<PY
import json
for c in ("c3","c2","c1","c4","c5","c3_reference"):
    try:
        d=[json.loads(l) for l in open("gpurun_out/${tag}_bench_%s.json"%c) if l.startswith("{")][0]
        print(c, "step", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "kern", (d.get("roofline") or {}).get("kernel_ms"), "frac", (d.get("roofline") or {}).get("frac"), d.get("parity"), d.get("clocks"))
    except Exception as e: print(c, "failed", e)
PY
