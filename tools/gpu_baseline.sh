#!/bin/bash
# One GPU-box pass: parity tests, bench line, launch lists (seg + cls step).  Usage (under gpurun): bash tools/gpu_baseline.sh TAG
TAG=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "test rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_$TAG.log
python tools/profile_step.py > gpurun_out/plain_step_$TAG.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_seg.csv python tools/profile_step.py > gpurun_out/ncu_step_$TAG.log 2>&1
python tools/profile_step.py T1_fetal_planes > gpurun_out/plain_step_cls_$TAG.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_cls.csv python tools/profile_step.py T1_fetal_planes > gpurun_out/ncu_step_cls_$TAG.log 2>&1
tail -1 gpurun_out/plain_step_$TAG.log gpurun_out/plain_step_cls_$TAG.log
