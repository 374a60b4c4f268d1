"""Runs ONE kernel shape a few times (for ncu captures).  python tools/one_kernel.py linear <stage 1-4> <qkv|proj|fc1|fc2> <fwd|dgrad|wgrad> [reps]
python tools/one_kernel.py lnbwd|lnfwd|attnfwd|attnbwd <stage> [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mtus_b200 import kbench

dev = torch.device("cuda", 0)
what = sys.argv[1]
stage = int(sys.argv[2])
_, M, Cc = kbench.swin_b_stage_shapes(32)[stage - 1]
if what == "linear":
    name, direction = sys.argv[3], sys.argv[4]
    N, K, kind = [(n, k, kd) for nm, n, k, kd in kbench.linear_cases(M, Cc) if nm == name][0]
    t, fl, by = kbench.time_linear(M, N, K, kind, direction, dev)
    print(f"{name}.{direction} stage {stage}: {t * 1e6:.1f} us, {fl / t / 1e12:.1f} TFLOP/s, {by / t / 1e9:.0f} GB/s algorithmic")
elif what in ("lnfwd", "lnbwd"):
    t, by = (kbench.time_layernorm_fwd if what == "lnfwd" else kbench.time_layernorm_bwd)(M, Cc, dev)
    print(f"{what} stage {stage}: {t * 1e6:.1f} us, {by / t / 1e9:.0f} GB/s")
else:
    H = {1: 56, 2: 28, 3: 14, 4: 7}[stage]
    t, by, fl = kbench.time_window_attn(32, H, Cc, Cc // 32, 7, 3 if stage < 4 else 0, dev, backward=(what == "attnbwd"))
    print(f"{what} stage {stage}: {t * 1e6:.1f} us, {by / t / 1e9:.0f} GB/s, {fl / t / 1e12:.1f} attention-only TFLOP/s")
