import sys, os, ctypes as C
sys.path.insert(0, "/root/repo")
import torch, mtus_b200 as m
from mtus_b200 import _lib
def stats():
    h, ms, n = C.c_int64(), C.c_int64(), C.c_int64()
    _lib.lib().mtus_graph_cache_stats(C.byref(h), C.byref(ms), C.byref(n))
    return h.value, ms.value, n.value
torch.manual_seed(0)
enc = m.SwinTransformerEncoder("swin_t", pretrained=False, img_size=224, precision="bf16", drop_path_rate=0.0).cuda().train()
x = torch.randn(2, 3, 224, 224).cuda()
gs = None
for it in range(4):
    for p in enc.parameters():
        p.grad = None
    feats = enc(x)
    print("after fwd", it, stats(), [hex(f.data_ptr()) for f in feats][:1])
    if gs is None:
        gs = [torch.randn_like(f) for f in feats]
    torch.autograd.backward(feats, gs)
    print("after bwd", it, stats(), hex(enc.model._last_flat_grad.data_ptr()))
