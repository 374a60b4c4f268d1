#!/bin/bash
# ncu --set full captures of the top kernels inside one real training step (classification task = encoder only).
# launch-skip values pick a stage-3 (C=512, M=6272) instance of each kernel.
TAG=${1:-r1}
for spec in "lnv2_bwd:10:ln_bwd" "window_attn_mma_bwd:6:attn_bwd" "window_attn_mma_fwd:8:attn_fwd" "lnv2_fwd:12:ln_fwd"; do
  IFS=: read k skip name <<< "$spec"
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k --launch-skip $skip --launch-count 1 \
      -o gpurun_out/prof_${name}_$TAG -f python tools/profile_step.py T1_fetal_planes > gpurun_out/ncu_${name}_$TAG.log 2>&1
done
# GEMMs: forward fc1 (+GELU, two outputs), forward fc2 (fp32 out + residual), dgrad, wgrad at stage 3
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc2 --launch-skip 40 --launch-count 8 \
    -o gpurun_out/prof_gemm_fwd_$TAG -f python tools/profile_step.py T1_fetal_planes > gpurun_out/ncu_gemm_fwd_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc2 --launch-skip 140 --launch-count 8 \
    -o gpurun_out/prof_gemm_bwd_$TAG -f python tools/profile_step.py T1_fetal_planes > gpurun_out/ncu_gemm_bwd_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep
