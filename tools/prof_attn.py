"""Runs window attention fwd + bwd at one Swin-B stage shape (default stage 3: 14x14, C=512, shifted) a few times:
the target of the `ncu --set full` capture of the attention kernels.   python tools/prof_attn.py [res C heads shift]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mtus_b200 import ops

dev = "cuda"
res, C, heads, shift = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (14, 512, 16, 3)
B = 32
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
qkv = torch.randn(B, res, res, 3 * C, device=dev).bfloat16()
tab = torch.randn(169, heads, device=dev) * 0.1
bias = torch.zeros(3 * C, device=dev)
out, lse = ops.window_attn_fwd(qkv, tab, bias, heads, 7, shift, return_lse=True)
dout = torch.randn_like(out)
for rep in range(2):
    flush.add_(1.0)
    out, lse = ops.window_attn_fwd(qkv, tab, bias, heads, 7, shift, return_lse=True)
    flush.add_(1.0)
    ops.window_attn_bwd(dout, qkv, out, tab, bias, heads, 7, shift, lse=lse, with_colsum=True)
torch.cuda.synchronize()
print("ok")
