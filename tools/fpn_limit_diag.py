"""Is the bf16 gradient-cosine floor on FPN-path tasks a property of the FPN kernels or of bf16 arithmetic upstream?

Experiment (swin_b, 224x224, batch 8, random init, loss = mean(out^2), segmentation task): feed the EXACT fp32 oracle
decoder + head with the bf16 kernel path's encoder features (cast to fp32, gradients flowing back into the bf16 kernel
encoder) and compare every parameter gradient with the all-fp32 oracle.  If the cosines of this hybrid are already below
0.999, no decoder precision (fp32, tf32, split-bf16) can reach the target: the perturbation that flips GroupNorm->ReLU
masks comes from the encoder's bf16 forward.   python tools/fpn_limit_diag.py
"""
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
cfg.config["model"]["decoder"]["dropout"] = 0.0
torch.manual_seed(0)
oracle = OracleMultiTaskModel(cfg, drop_path_rate=0.0).cuda().eval()
model = m.build_model(cfg, precision="bf16").cuda().eval()
model.load_state_dict(oracle.state_dict())
hybrid_tail = OracleMultiTaskModel(cfg, drop_path_rate=0.0).cuda().eval()
hybrid_tail.load_state_dict(oracle.state_dict())
x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(5)).cuda()
tid = "T2B_adult_liver_segment_5"


def cosines(named_a, named_b):
    rows = []
    for k, ga in named_a.items():
        gb = named_b.get(k)
        if ga is None or gb is None or gb.norm() == 0:
            continue
        rows.append((torch.nn.functional.cosine_similarity(ga.float().flatten(), gb.float().flatten(), dim=0).item(), k))
    rows.sort()
    return rows


# all-fp32 oracle
oracle.zero_grad(set_to_none=True)
oracle(x, tid).square().mean().backward()
ref = {k: p.grad for k, p in oracle.named_parameters()}

# full bf16 kernel path
model.zero_grad(set_to_none=True)
model(x, tid).float().square().mean().backward()
full = {k: p.grad for k, p in model.named_parameters()}

# hybrid: bf16 kernel encoder -> fp32 oracle decoder + head
model.zero_grad(set_to_none=True)
hybrid_tail.zero_grad(set_to_none=True)
feats = [f.float().contiguous() for f in model.encoder(x)]
hybrid_tail.heads[tid](hybrid_tail.fpn_decoder_seg(feats)).square().mean().backward()
hyb = {k: p.grad for k, p in model.named_parameters() if k.startswith("encoder.")}
hyb.update({k: p.grad for k, p in hybrid_tail.named_parameters() if not k.startswith("encoder.")})

# straight-through: EVERYTHING exact fp32 (oracle encoder, decoder, head, backward) -- only the VALUES of the four features
# are replaced by the bf16 kernel encoder's in the forward pass; gradients flow through the oracle encoder unperturbed
st_model = OracleMultiTaskModel(cfg, drop_path_rate=0.0).cuda().eval()
st_model.load_state_dict(oracle.state_dict())
with torch.no_grad():
    fk = [f.float().contiguous() for f in model.encoder(x)]
fo = st_model.encoder(x)
mixed = [o + (k - o).detach() for o, k in zip(fo, fk)]
st_model.heads[tid](st_model.fpn_decoder_seg(mixed)).square().mean().backward()
stt = {k: p.grad for k, p in st_model.named_parameters()}
feat_err = [((k - o).norm() / o.norm()).item() for o, k in zip(fo, fk)]
print("relative L2 error of the bf16 encoder's four features vs the fp32 oracle:", ", ".join(f"{e:.2e}" for e in feat_err))

for label, got in (("bf16 kernel path (encoder + FPN + head)", full), ("bf16 kernel encoder + EXACT fp32 decoder/head", hyb),
                   ("ALL fp32, only the forward VALUES of the features taken from the bf16 encoder (straight-through)", stt)):
    rows = cosines(got, ref)
    below = [r for r in rows if r[0] < 0.999]
    dec = [r for r in rows if r[1].startswith("fpn_decoder")]
    enc = [r for r in rows if r[1].startswith("encoder.")]
    print(f"{label}: {len(rows)} tensors, {len(below)} below 0.999, min {rows[0][0]:.5f} ({rows[0][1]})")
    print(f"    decoder tensors: min {min(dec)[0]:.5f}   encoder tensors: min {min(enc)[0]:.5f}")
    for c, k in below[:12]:
        print(f"      {c:.5f}  {k}")
