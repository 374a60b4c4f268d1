"""Eager vs CUDA-graph replay of the encoder executors (mtus_swin_forward / mtus_swin_backward), Swin-B, B=32:
how much of the step is launch gaps.  python tools/graph_probe.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mtus_b200 as m
from mtus_b200 import _lib

dev = torch.device("cuda", 0)
torch.manual_seed(0)
enc = m.encoders.SwinTransformerEncoder("swin_b", pretrained=False, img_size=224, precision="bf16").to(dev).train()
core = enc.model
L = _lib.lib()
B = 32
x = torch.randn(B, 3, 224, 224, device=dev)
cfg = core._cfg(B, True)
flat = core.flat_params()
ws = torch.empty(L.mtus_swin_workspace_bytes(C.byref(cfg)), dtype=torch.uint8, device=dev)
lp = flat.to(torch.bfloat16)
dp = core._droppath_scales(B, dev)
feats = [torch.empty(B, 128 * 2 ** i, 56 // 2 ** i, 56 // 2 ** i, dtype=torch.bfloat16, device=dev) for i in range(4)]
dfeats = [torch.randn_like(f) for f in feats]
grad = torch.zeros_like(flat)


def fwd():
    _lib.check(L.mtus_swin_forward(C.byref(cfg), _lib.ptr(x), 1, _lib.ptr(flat), _lib.ptr(lp), _lib.ptr(dp), _lib.ptr(ws),
                                   _lib.ptr_array(feats), 0, 0, _lib.stream_ptr()), "fwd")


def bwd():
    _lib.check(L.mtus_swin_backward(C.byref(cfg), _lib.ptr(flat), _lib.ptr(lp), _lib.ptr(dp), _lib.ptr(ws), _lib.ptr_array(dfeats), 0, 0,
                                    _lib.ptr(grad), 4, 0, _lib.stream_ptr()), "bwd")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for name, fn in (("forward", fwd), ("backward", bwd)):
    t_eager = timeit(fn)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    t_graph = timeit(g.replay)
    print(f"{name}: eager {t_eager:.3f} ms, graph replay {t_graph:.3f} ms")
