"""Per-shape roofline table of the tensor-core GEMM engine: Swin-B (batch 32, 224x224), 4 stages x {qkv, proj, fc1, fc2} x
{fwd, dgrad, wgrad}, timed host-free (CUDA graphs over operand rings > 4x L2, mtus_b200/kbench.py) against the measured
sustained bf16 peak.   python tools/roofline_table.py [--stages 3,4] > profiles/rN_gemm_roofline_table.txt"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mtus_b200 import kbench

ap = argparse.ArgumentParser()
ap.add_argument("--stages", default="1,2,3,4")
ap.add_argument("--batch", type=int, default=32)
args = ap.parse_args()
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    src = "measured sustained"
except Exception:
    peak, src = 1400.0, "fallback"
dev = torch.device("cuda", 0)
depth = {1: 2, 2: 2, 3: 18, 4: 2}
print(f"gemm_tc2_kernel, Swin-B batch {args.batch}: us per launch | TFLOP/s | fraction of {peak:.0f} TFLOP/s ({src})")
print(f"{'stage':5s} {'M':>7s} {'layer':5s} | {'fwd':>24s} | {'dgrad':>24s} | {'wgrad':>24s}")
tot_fl = tot_t = 0.0
for stage, M, Cc in kbench.swin_b_stage_shapes(args.batch):
    if str(stage) not in args.stages.split(","):
        continue
    for name, N, K, kind in kbench.linear_cases(M, Cc):
        cells = []
        for direction in ("fwd", "dgrad", "wgrad"):
            t, fl, by = kbench.time_linear(M, N, K, kind, direction, dev)
            cells.append(f"{t * 1e6:7.1f} {fl / t / 1e12:7.1f} {fl / t / 1e12 / peak:6.3f}")
            tot_fl += fl * depth[stage]
            tot_t += t * depth[stage]
        print(f"{stage:5d} {M:7d} {name:5s} | " + " | ".join(f"{c:>24s}" for c in cells), flush=True)
print(f"block-weighted (depths 2/2/18/2) total: {tot_t * 1e3:.3f} ms of GEMM per training step, {tot_fl / tot_t / 1e12:.1f} TFLOP/s = "
      f"{tot_fl / tot_t / 1e12 / peak:.3f} of peak")
