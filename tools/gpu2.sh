#!/bin/bash
# multi-GPU pass: the whole GPU suite (multi-device tests included) and the torchrun bench line with its sharded sub-records.
tag=${1:-r2f}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "bench rc=$?"; tail -5 gpurun_out/${tag}_bench_n$N.err; cut -c1-300 gpurun_out/${tag}_bench_n$N.json
