"""Where does the bf16 gradient error of a segmentation step enter?  Cosine (bf16 kernel path vs fp32 oracle) of the
gradient w.r.t. the FPN output (head backward), the four encoder features (FPN backward) and the output itself."""
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
torch.manual_seed(0)
oracle = OracleMultiTaskModel(cfg, drop_path_rate=0.0).cuda().eval()
model = m.build_model(cfg, precision="bf16").cuda().eval()
model.load_state_dict(oracle.state_dict())
x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(5)).cuda()
tid = "T2B_adult_liver_segment_5"


def tap(net, store):
    def fpn_hook(mod, i, o):
        store["fpn"] = o.detach().float().clone()
        o.register_hook(lambda g: store.__setitem__("dfpn", g.detach().float().clone()))

    def enc_hook(mod, i, o):
        for k, f in enumerate(o):
            f.register_hook(lambda g, k=k: store.__setitem__(f"dfeat{k}", g.detach().float().clone()))
        store["feats"] = [f.detach().float().clone() for f in o]

    net.fpn_decoder_seg.register_forward_hook(fpn_hook)
    net.encoder.register_forward_hook(enc_hook)


so, sm = {}, {}
tap(oracle, so); tap(model, sm)
yo, ym = oracle(x, tid), model(x, tid)
yo.square().mean().backward()
ym.float().square().mean().backward()
cos = lambda a, b: torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()
print("output          cos", cos(ym.float(), yo))
print("fpn output      cos", cos(sm["fpn"], so["fpn"]))
for k in range(4):
    print(f"feature {k}       cos", cos(sm["feats"][k], so["feats"][k]))
print("d fpn output    cos", cos(sm["dfpn"], so["dfpn"]))
for k in range(4):
    print(f"d feature {k}     cos", cos(sm[f"dfeat{k}"], so[f"dfeat{k}"]))
# isolate the FPN backward: feed the ORACLE's gradient of the FPN output through the kernel FPN backward
model.zero_grad(set_to_none=True)
sm2 = {}
feats = model.encoder(x)
for k, f in enumerate(feats):
    f.register_hook(lambda g, k=k: sm2.__setitem__(f"dfeat{k}", g.detach().float().clone()))
out = model.fpn_decoder_seg(feats)
out.backward(so["dfpn"].to(out.dtype))
for k in range(4):
    print(f"d feature {k} with the oracle's d fpn output: cos", cos(sm2[f"dfeat{k}"], so[f"dfeat{k}"]))

# same, but the FPN takes the oracle's gradient in fp32 (output_dtype='fp32' decoder): is the loss of precision in the
# bf16 rounding of the incoming gradient or inside the decoder's backward?
src = model.fpn_decoder_seg
dec = m.FPNDecoder(model.encoder.out_channels, encoder_depth=4, pyramid_channels=src.pyramid_channels,
                   segmentation_channels=src.segmentation_channels, dropout=0.0, merge_policy=src.merge_policy,
                   precision="bf16", output_dtype="fp32").cuda().eval()
dec.load_state_dict(src.state_dict())
sm3 = {}
feats = model.encoder(x)
for k, f in enumerate(feats):
    f.register_hook(lambda g, k=k: sm3.__setitem__(f"dfeat{k}", g.detach().float().clone()))
out = dec(feats)
out.backward(so["dfpn"].to(out.dtype))
for k in range(4):
    print(f"d feature {k} with the oracle's d fpn output in fp32: cos", cos(sm3[f"dfeat{k}"], so[f"dfeat{k}"]))
