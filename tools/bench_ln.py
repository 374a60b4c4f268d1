"""Times the mixed-precision LayerNorm family at the Swin-B stage shapes (B=32); CUDA events, L2 flushed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mtus_b200 import ops

dev = "cuda"
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


for rows, C in ((32 * 3136, 128), (32 * 784, 256), (32 * 196, 512), (32 * 49, 1024), (32 * 784, 512), (32 * 196, 1024), (32 * 49, 2048)):
    x = torch.randn(rows, C, device=dev)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    y, mean, rstd = ops.layernorm_fwd_mixed(x, g, b, torch.bfloat16)
    dy = torch.randn(rows, C, device=dev).bfloat16()
    dres = torch.randn(rows, C, device=dev)
    tf = timeit(lambda: ops.layernorm_fwd_mixed(x, g, b, torch.bfloat16))
    tb = timeit(lambda: ops.layernorm_bwd_mixed(dy, x, g, mean, rstd, dres, lp_dtype=torch.bfloat16))
    bf, bb = rows * C * 6.0, rows * C * (2 + 4 + 4 + 4 + 2.0)
    print(f"[{rows:6d},{C:4d}] fwd {tf * 1e6:7.1f} us {bf / tf / 1e9:7.1f} GB/s | bwd {tb * 1e6:7.1f} us {bb / tb / 1e9:7.1f} GB/s")
