#!/bin/bash
# N-GPU scaling pass as the driver runs it: the default bench line at N (with its sharded C4 / C5 / one-process sub-records).
tag=${1:-r2l}
N=${2:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("C3 n$N step", d["ms_per_step"], "wall", d["wall_ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["clocks"])
        for k,v in d.get("sharded",{}).items():
            print(k, {kk: v.get(kk) for kk in ("ms_per_step","wall_ms_per_step","parity","exchange_bytes_per_step","matches_per_step","error","value")})
            if "roofline" in v: print("   kernel_ms", v["roofline"]["kernel_ms"], v["roofline"].get("kernel_ms_max_over_ranks"))
PY
