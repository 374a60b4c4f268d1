#!/bin/bash
# One short GPU-box pass over the final build: parity tests, bench line (default flags), launch lists of one classification and
# one segmentation training step (ncu time-only pass, summarised on the box).  Usage (under gpurun): bash tools/final_validation.sh TAG
TAG=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/${TAG}_pytest_gpu.log | tail -2
python bench.py > gpurun_out/${TAG}_bench_1gpu.json.log 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; grep '^{' gpurun_out/${TAG}_bench_1gpu.json.log | head -c 300; echo
for spec in "cls:T1_fetal_planes" "seg:T2A_fetal_abdomen"; do
  IFS=: read name tid <<< "$spec"
  timeout 120 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/${TAG}_step_launches_${name}_b32.csv python tools/profile_step.py $tid > gpurun_out/${TAG}_ncu_step_${name}.log 2>&1
  echo "ncu $name rc=$?"
  python tools/summarize_launches.py gpurun_out/${TAG}_step_launches_${name}_b32.csv 60 > gpurun_out/${TAG}_step_launches_${name}_b32.txt 2>&1
  head -4 gpurun_out/${TAG}_step_launches_${name}_b32.txt
done
