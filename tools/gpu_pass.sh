#!/bin/bash
# single-GPU pass: parity suite, headline benches, launch lists
tag=${1:-pass}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/${tag}_pytest_gpu.log; tail -4 gpurun_out/${tag}_pytest_gpu.log
for c in C3 C2 C1 C4; do
  timeout 300 python bench.py --workload $c --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; echo "bench $c rc=$?"
done
python - <<PY
import json
for c in ("C3","C2","C1","C4"):
    try:
        d=[json.loads(l) for l in open("gpurun_out/${tag}_bench_%s.json"%c) if l.startswith("{")][0]
        print(c, "step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "kern", round(d["roofline"]["kernel_ms"],4), "launches/step", d["gpu_launches"]/d["steps"], d.get("parity"), d["clocks"])
    except Exception as e: print(c, "failed", e)
PY
if [ "$2" == "ncu" ]; then
for c in C3 C1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches_$c.csv python bench.py --workload $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_under_ncu_$c.log 2>&1
done
fi
