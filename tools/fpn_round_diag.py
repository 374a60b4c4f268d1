"""Which bf16 rounding point inside the FPN costs gradient cosine?  The fp32 oracle decoder is run on the oracle's own
features / output gradient with bf16 rounding injected at one class of points at a time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
import mtus_b200 as m
from oracle.model import OracleMultiTaskModel
from oracle.fpn import Conv3x3GNReLU

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


class RG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rf, rb):
        ctx.rb = rb
        return x.bfloat16().float() if rf else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (g.bfloat16().float() if ctx.rb else g), None, None


B = 8
cfg = m.swin_b_27task(batch_size=B)
torch.manual_seed(0)
oracle = OracleMultiTaskModel(cfg, drop_path_rate=0.0).cuda().eval()
x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(5)).cuda()
tid = "T2B_adult_liver_segment_5"
store = {}
oracle.fpn_decoder_seg.register_forward_hook(lambda mod, i, o: o.register_hook(lambda g: store.__setitem__("dfpn", g.detach().clone())) and None)
h = oracle.encoder.register_forward_hook(lambda mod, i, o: store.__setitem__("feats", [f.detach().clone() for f in o]))
oracle(x, tid).square().mean().backward()
dec = oracle.fpn_decoder_seg
dec._forward_hooks.clear()
feats, dfpn = store["feats"], store["dfpn"]
cos = lambda a, b: torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()


def run(mode):
    hs = []
    saved_w = {}
    convs = [mod for mod in dec.modules() if isinstance(mod, nn.Conv2d)]
    if mode in ("weights", "all"):
        for c in convs:
            saved_w[c] = c.weight.data.clone()
            c.weight.data = c.weight.data.bfloat16().float()
    for c in convs:
        if mode in ("conv_inputs", "all"):
            hs.append(c.register_forward_pre_hook(lambda mod, i: (RG.apply(i[0], True, False),)))
        if mode in ("grad_at_conv_inputs", "all"):
            hs.append(c.register_forward_pre_hook(lambda mod, i: (RG.apply(i[0], False, True),)))
        if mode in ("grad_at_conv_outputs", "all"):
            hs.append(c.register_forward_hook(lambda mod, i, o: RG.apply(o, False, True)))
    fs = [f.clone().bfloat16().float().requires_grad_(True) if mode in ("features", "all") else f.clone().requires_grad_(True) for f in feats]
    out = dec(fs)
    out.backward(dfpn)
    for hh in hs:
        hh.remove()
    for c, w in saved_w.items():
        c.weight.data = w
    return [f.grad.clone() for f in fs]


ref = run("none")
for mode in ("features", "weights", "conv_inputs", "grad_at_conv_inputs", "grad_at_conv_outputs", "all"):
    g = run(mode)
    print(f"{mode:22s}", " ".join(f"{cos(a, b):.5f}" for a, b in zip(g, ref)))
