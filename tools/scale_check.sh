#!/bin/bash
# Multi-GPU pass on one box: tools/scale_check.sh N  -> gpurun_out/scale_n<N>_*.json (default workload, C4, C5)
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --steps 10 --warmup 3 > gpurun_out/scale_n${N}_C3.json 2> gpurun_out/scale_n${N}_C3.err; echo "C3 rc=$?"
run --workload C4 --steps 5 --warmup 3 > gpurun_out/scale_n${N}_C4.json 2> gpurun_out/scale_n${N}_C4.err; echo "C4 rc=$?"
run --workload C5 --steps 3 --warmup 3 > gpurun_out/scale_n${N}_C5.json 2> gpurun_out/scale_n${N}_C5.err; echo "C5 rc=$?"
for c in C3 C4 C5; do tail -c 1500 gpurun_out/scale_n${N}_$c.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$c n=',d['n_gpus'],'ms',round(d['ms_per_step'],3),'value %.3e'%d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), d['config'].get('parallelism'))
" 2>/dev/null || tail -3 gpurun_out/scale_n${N}_$c.err; done
