#!/bin/bash
# bench.py --workload W at N GPUs (the driver's launch line).  Usage (under gpurun --gpus N): bash tools/dp_scale_workload.sh N W TAG
N=${1:-2}; W=${2:-swin_l_384}; TAG=${3:-r2}
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --workload $W --steps 20 --warmup 5 > gpurun_out/bench_${W}_${N}gpu_$TAG.log 2>&1
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --workload $W --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${W}_${N}gpu_$TAG.log 2>&1
fi
echo "rc=$?"; tail -1 gpurun_out/bench_${W}_${N}gpu_$TAG.log | cut -c1-260
