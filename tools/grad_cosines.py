"""Per-parameter gradient cosine of the bf16 kernel path against the fp32 oracle at the headline geometry (swin_b, 27 heads,
224x224, batch 8): how many tensors are below 0.999 and which.   python tools/grad_cosines.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mtus_b200 as m
from oracle.model import OracleMultiTaskModel

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B = 8
cfg = m.swin_b_27task(batch_size=B)
cfg.config["model"]["decoder"]["dropout"] = 0.0 if "decoder" in cfg.config.get("model", {}) else 0.0
torch.manual_seed(0)
oracle = OracleMultiTaskModel(cfg, drop_path_rate=0.0).cuda().eval()
model = m.build_model(cfg, precision="bf16").cuda().eval()
model.load_state_dict(oracle.state_dict())
x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(5)).cuda()
for tid in ("T2B_adult_liver_segment_5", "T3C_thyroid_nodule"):
    oracle.zero_grad(set_to_none=True)
    model.zero_grad(set_to_none=True)
    yo, ym = oracle(x, tid), model(x, tid)
    rel = ((ym.float() - yo).abs().max() / yo.abs().max()).item()
    yo.square().mean().backward()
    ym.float().square().mean().backward()
    po = dict(oracle.named_parameters())
    rows = []
    for name, p in model.named_parameters():
        go = po[name].grad
        if go is None or p.grad is None or go.norm() == 0:
            continue
        c = torch.nn.functional.cosine_similarity(p.grad.float().flatten(), go.flatten(), dim=0).item()
        rows.append((c, name, go.numel()))
    rows.sort()
    below = [r for r in rows if r[0] < 0.999]
    print(f"{tid}: output rel err {rel:.4f}; {len(rows)} tensors, {len(below)} below 0.999, min {rows[0][0]:.5f}")
    for c, name, n in below[:40]:
        print(f"   {c:.5f}  {name}  ({n} elements)")
