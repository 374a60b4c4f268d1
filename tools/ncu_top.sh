#!/bin/bash
# ncu --set full captures of the top kernels inside one real training step (classification task = encoder only),
# summarised on the box with tools/ncu_summary.py (the .ncu-rep files are too big to ship back; set KEEP_REP=1 to keep them).
# launch-skip values pick a stage-3 (C=512, M=6272) instance of each kernel.
TAG=${1:-r1}
export MTUS_GRAPHS=0   # per-kernel capture: profile the plain schedule (graph nodes would need --graph-profiling node)
for spec in "lnv2_bwd:10:ln_bwd:1" "lnv2_fwd:12:ln_fwd:1" "window_attn_mma_bwd:6:attn_bwd:1" "window_attn_mma_fwd:8:attn_fwd:1" "gemm_tc2:40:gemm_fwd:8" "gemm_tc2:140:gemm_bwd:8"; do
  IFS=: read k skip name count <<< "$spec"
  ncu --set full --clock-control none --profile-from-start off -k regex:$k --launch-skip $skip --launch-count $count \
      -o gpurun_out/prof_${name}_$TAG -f python tools/profile_step.py T1_fetal_planes > gpurun_out/ncu_${name}_$TAG.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${name}_$TAG.ncu-rep > gpurun_out/ncu_summary_${name}_$TAG.txt 2>&1
  [ -z "$KEEP_REP" ] && rm -f gpurun_out/prof_${name}_$TAG.ncu-rep
done
ls -la gpurun_out/ | head -40
