"""Times the GEMM shapes of one Swin-B training step (B=32, 224x224) through the C-ABI; CUDA events, L2 flushed.

    python tools/bench_gemm.py            (MTUS_GEMM=tc1 selects the one-tile-per-CTA engine for A/B runs)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mtus_b200 import ops, _lib

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
    B = 32
    stages = [(B * 3136, 128, 2), (B * 784, 256, 2), (B * 196, 512, 18), (B * 49, 1024, 2)]
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    totf = 0.0
    print(f"{'shape':34s} {'fwd us':>8} {'TF/s':>7} {'dgrad us':>9} {'TF/s':>7} {'wgrad us':>9} {'TF/s':>7}")
    for M, C, depth in stages:
        for name, K, N, kind in (("qkv", C, 3 * C, "bias"), ("proj", C, C, "res"), ("fc1", C, 4 * C, "gelu"), ("fc2", 4 * C, C, "res")):
            x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
            w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
            b = torch.zeros(N, device=dev)
            res = torch.randn(M, N, device=dev).bfloat16()
            dy = torch.randn(M, N, device=dev).bfloat16()
            pre = torch.randn(M, K, device=dev).bfloat16()
            if kind == "bias":
                f = lambda: ops.linear_fwd(x, w, b, backend=2)
            elif kind == "res":
                f = lambda: ops.linear_fwd(x, w, b, res=res, backend=2)
            else:
                f = lambda: ops.linear_fwd(x, w, b, gelu=True, backend=2)
            tf = timeit(f)
            td = timeit((lambda: ops.linear_dgrad(dy, w, gelu_pre=pre, backend=2)) if name == "fc2" else (lambda: ops.linear_dgrad(dy, w, backend=2)))
            tw = timeit(lambda: ops.linear_wgrad(dy, x, with_bias=False, backend=2))
            fl = 2.0 * M * K * N
            print(f"{name:5s} M{M:6d} K{K:5d} N{N:5d} x{depth:2d}   {tf * 1e6:8.1f} {fl / tf / 1e12:7.1f} {td * 1e6:9.1f} {fl / td / 1e12:7.1f} {tw * 1e6:9.1f} {fl / tw / 1e12:7.1f}")
            tot["fwd"] += tf * depth; tot["dgrad"] += td * depth; tot["wgrad"] += tw * depth
            totf += fl * depth
    for k, v in tot.items():
        print(f"total {k}: {v * 1e3:.3f} ms  ({totf / v / 1e12:.1f} TFLOP/s)")
    # FPN convs (B=32): 256->128 @56, 128->128 @28
    for (H, Cin, Cout) in ((56, 256, 128), (28, 256, 128), (28, 128, 128), (14, 256, 128)):
        x = (torch.randn(B, H, H, Cin, device=dev) * 0.5).bfloat16()
        wt = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
        wf, wd = ops.conv3x3_repack(wt, torch.bfloat16)
        dy = torch.randn(B, H, H, Cout, device=dev).bfloat16()
        fl = 2.0 * B * H * H * 9 * Cin * Cout
        tf = timeit(lambda: ops.conv3x3_fwd(x, wf, backend=2))
        td = timeit(lambda: ops.conv3x3_dgrad(dy, wd, backend=2))
        tw = timeit(lambda: ops.conv3x3_wgrad(dy, x, backend=0))
        print(f"conv3x3 {H}x{H} {Cin}->{Cout}: fwd {tf * 1e6:.1f} us {fl / tf / 1e12:.1f} TF/s | dgrad {td * 1e6:.1f} us {fl / td / 1e12:.1f} | wgrad {tw * 1e6:.1f} us {fl / tw / 1e12:.1f}")


if __name__ == "__main__":
    main()
