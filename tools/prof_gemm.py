"""Runs the Swin-B stage-3 forward GEMMs (qkv+bias, proj+residual, fc1+GELU, fc2+residual; M = 6272) a few times:
the target of the `ncu --set full` capture of the dominant kernel (gemm_tc2_kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mtus_b200 import ops

dev = "cuda"
M, C = 32 * 196, 512
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for rep in range(3):
    for name, K, N, kind in (("qkv", C, 3 * C, "bias"), ("proj", C, C, "res"), ("fc1", C, 4 * C, "gelu"), ("fc2", 4 * C, C, "res")):
        x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        b = torch.zeros(N, device=dev)
        res = torch.randn(M, N, device=dev).bfloat16()
        flush.add_(1.0)
        if kind == "bias":
            ops.linear_fwd(x, w, b, backend=2)
        elif kind == "res":
            ops.linear_fwd(x, w, b, res=res, backend=2)
        else:
            ops.linear_fwd(x, w, b, gelu=True, backend=2)
torch.cuda.synchronize()
print("ok")
