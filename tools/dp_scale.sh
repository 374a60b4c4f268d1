#!/bin/bash
# bench.py at N GPUs (N = $1), the driver's launch line
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu_r1h.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/bench_${N}gpu_r1h.log | cut -c1-400
