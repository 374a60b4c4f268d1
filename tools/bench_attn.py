"""Times window attention fwd / bwd at the four Swin-B stage shapes (B=32, 224x224); CUDA events, L2 flushed."""
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


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    tot_f = tot_b = 0.0
    for (res, C, heads, depth) in ((56, 128, 4, 2), (28, 256, 8, 2), (14, 512, 16, 18), (7, 1024, 32, 2)):
        qkv = torch.randn(B, res, res, 3 * C, device=dev).bfloat16()
        tab = torch.randn(169, heads, device=dev) * 0.1
        bias = torch.zeros(3 * C, device=dev)
        for shift in ((0, 3) if res > 7 else (0,)):
            out, lse = ops.window_attn_fwd(qkv, tab, bias, heads, 7, shift, return_lse=True)
            dout = torch.randn_like(out)
            tf = timeit(lambda: ops.window_attn_fwd(qkv, tab, bias, heads, 7, shift))
            tb = timeit(lambda: ops.window_attn_bwd(dout, qkv, out, tab, bias, heads, 7, shift, lse=lse, with_colsum=True))
            tokens = B * res * res
            byf, byb = 4.0 * tokens * C * 2, 8.0 * tokens * C * 2
            fl = 4.0 * 49 * 49 * 32 * (tokens / 49) * heads
            print(f"res {res:2d} C {C:4d} shift {shift}: fwd {tf * 1e6:7.1f} us {byf / tf / 1e9:7.1f} GB/s {fl / tf / 1e12:6.2f} TF/s | "
                  f"bwd {tb * 1e6:7.1f} us {byb / tb / 1e9:7.1f} GB/s {2.5 * fl / tb / 1e12:6.2f} TF/s")
            n = depth / (2 if res > 7 else 1)
            tot_f += tf * n; tot_b += tb * n
    print(f"per step (24 blocks): fwd {tot_f * 1e3:.3f} ms, bwd {tot_b * 1e3:.3f} ms")


if __name__ == "__main__":
    main()
