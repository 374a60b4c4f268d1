#!/bin/bash
# ncu --set full of the bandwidth-bound kernels reworked in round 2 (session 3), one launch each out of tools/hbm_kernels.py
# (the eager warm-up launches in front of the graph), summarised with tools/ncu_summary.py.  Usage (under gpurun): bash tools/gpu_hbm_profiles.sh TAG
TAG=${1:-r2}
mkdir -p gpurun_out
for spec in "gn_fused_bwd:gn_bwd relu [32,56:gn_bwd" "colsum_kernel:colsum [100352:colsum" "lnx_bwd_kernel:merge_ln_bwd [32,14:lnx_bwd_wide" "lnv2_bwd_kernel:ln_bwd s3:ln_bwd_s3"; do
  IFS=: read k filt name <<< "$spec"
  ncu --set full --clock-control none -k regex:$k --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${name}_$TAG -f \
      python tools/hbm_kernels.py "$filt" > gpurun_out/ncu_${name}_$TAG.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${name}_$TAG.ncu-rep > gpurun_out/ncu_summary_${name}_$TAG.txt 2>&1
  rm -f gpurun_out/prof_${name}_$TAG.ncu-rep
done
cat gpurun_out/ncu_summary_*_$TAG.txt | head -120
