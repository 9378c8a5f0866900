N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload C4 --steps 10 --warmup 3 > gpurun_out/scale2_n8_C4.json 2> gpurun_out/scale2_n8_C4.err; echo "C4 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --workload C4 --steps 10 --warmup 3 > gpurun_out/scale2_n4_C4.json 2> gpurun_out/scale2_n4_C4.err; echo "C4 n4 rc=$?"
grep "^{" gpurun_out/scale2_n8_C4.json | cut -c1-400
grep "^{" gpurun_out/scale2_n4_C4.json | cut -c1-400
