"""Swin-B stage-1 GEMMs (M = 100352, C = 128): HBM-bound shapes (K = 128): qkv+bias, proj (fp32 stream), fc1+GELU, fc2 and the
fc2 dgrad (GELU' epilogue).  Target of an ncu --set full capture."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mtus_b200 import ops, _lib

dev = "cuda"
M, C = 32 * 3136, 128
x = (torch.randn(M, C, device=dev) * 0.5).bfloat16()
a4 = (torch.randn(M, 4 * C, device=dev) * 0.5).bfloat16()
res = torch.randn(M, C, device=dev)
w_qkv = (torch.randn(3 * C, C, device=dev) * 0.02).bfloat16()
w_fc1 = (torch.randn(4 * C, C, device=dev) * 0.02).bfloat16()
w_fc2 = (torch.randn(C, 4 * C, device=dev) * 0.02).bfloat16()
b3, b4, b1 = torch.zeros(3 * C, device=dev), torch.zeros(4 * C, device=dev), torch.zeros(C, device=dev)
for rep in range(2):
    ops.linear_fwd(x, w_qkv, b3, backend=_lib.BACKEND_TCGEN05)
    y, h = ops.linear_fwd(x, w_fc1, b4, gelu=True, backend=_lib.BACKEND_TCGEN05)
    ops.linear_fwd_stream(a4, w_fc2, b1, res=res, backend=_lib.BACKEND_TCGEN05)
    ops.linear_dgrad(x, w_fc2, gelu_pre=h, backend=_lib.BACKEND_TCGEN05)
torch.cuda.synchronize()
print("ok")
