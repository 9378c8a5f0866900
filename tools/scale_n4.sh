N=4
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --steps 10 --warmup 3 > gpurun_out/scale_n4_C3.json 2> gpurun_out/scale_n4_C3.err; echo "C3 rc=$?"
run --workload C5 --steps 3 --warmup 3 > gpurun_out/scale_n4_C5.json 2> gpurun_out/scale_n4_C5.err; echo "C5 rc=$?"
grep "^{" gpurun_out/scale_n4_C3.json | cut -c1-220
grep "^{" gpurun_out/scale_n4_C5.json | cut -c1-220
