"""Generates tests/golden/*.pt by running the REFERENCE's own model code.

Run in the authoring container only (needs /root/reference):  python tests/golden/make_golden.py

The reference's ``code/models/{encoders,decoders,multitask_model,heads}.py`` are imported unmodified; the two
third-party packages they delegate to (timm, segmentation_models_pytorch -- not installable here) are provided by
``oracle.shims``.  Weights are NOT stored (9-110 MB): every case re-creates them from ``torch.manual_seed(seed)``
through ``OracleMultiTaskModel`` (same state-dict keys, loaded into the reference model with strict=True) and the
fixture keeps per-tensor checksums so a drifting RNG is detected instead of silently comparing different weights.
Stored per case: input seed, encoder features, decoder/head outputs (sub-sampled when large) and the gradient of
``out.float().square().mean()`` (per-parameter L2 norms + a few full small tensors).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = {
    # BASELINE.json configs[0]: swin_t + shared FPN, one segmentation task, 224x224, batch 4, fp32 fwd+bwd
    "config1_swin_t_224": dict(encoder="swin_t", img=224, batch=4, separate=False,
                               tasks=["T2A_fetal_abdomen"], sub=(8, 4)),
    # small 4-task-type case (seg / cls / det / reg routing), separate FPNs like swin_b.yaml
    "micro_64_4types": dict(encoder="swin_micro_patch4_window7_test", img=64, batch=2, separate=True,
                            tasks=["T2A_fetal_abdomen", "T1_fetal_planes", "T4A_fetal_brain", "T5_fetal_femur"],
                            sub=(4, 1)),
}
FULL_GRADS = ("relative_position_bias_table", "norm1.weight", "patch_embed.proj.bias", "p5.bias", "block.1.weight")


def build_case(name, spec):
    import mtus_b200 as m
    from oracle.model import OracleMultiTaskModel
    cfg = m.make_config(spec["encoder"], spec["img"], spec["batch"], separate_fpn=spec["separate"], dropout=0.0,
                        mixed_precision=False, tasks=[t for t in m.tasks_27() if t["task_id"] in spec["tasks"]])
    torch.manual_seed(0)
    oracle = OracleMultiTaskModel(cfg, drop_path_rate=0.0).eval()
    x = torch.randn(spec["batch"], 3, spec["img"], spec["img"], generator=torch.Generator().manual_seed(1))
    return cfg, oracle, x


def subsample(t, sub):
    cs, ss = sub
    if t.dim() == 4 and t.numel() > 65536:
        return t[:, ::cs, ::ss, ::ss].contiguous()
    return t.contiguous()


def main():
    from oracle import shims
    models, _, _ = shims.import_reference_models("/root/reference")
    for name, spec in CASES.items():
        cfg, oracle, x = build_case(name, spec)
        ref = models.build_model(cfg).eval()              # the reference's MultiTaskModel (multitask_model.py:346)
        ref.load_state_dict(oracle.state_dict(), strict=True)
        fx = {"spec": spec, "weight_checksums": {k: (float(v.double().sum()), float(v.double().abs().sum()))
                                                 for k, v in oracle.state_dict().items() if v.is_floating_point()}}
        with torch.no_grad():
            feats = ref.encoder(x)
        fx["features"] = [subsample(f, spec["sub"]) for f in feats]
        fx["feature_sums"] = [(float(f.double().sum()), float(f.double().abs().sum())) for f in feats]
        fx["outputs"], fx["output_sums"], fx["grad_norms"], fx["grads"] = {}, {}, {}, {}
        for tid in spec["tasks"]:
            ref.zero_grad(set_to_none=True)
            out = ref(x, tid)
            out.float().square().mean().backward()
            fx["outputs"][tid] = subsample(out.detach(), spec["sub"])
            fx["output_sums"][tid] = (float(out.double().sum()), float(out.double().abs().sum()))
            fx["grad_norms"][tid] = {k: float(p.grad.double().norm()) for k, p in ref.named_parameters() if p.grad is not None}
            fx["grads"][tid] = {k: p.grad.clone() for k, p in ref.named_parameters()
                                if p.grad is not None and k.endswith(FULL_GRADS) and p.numel() <= 4096}
        path = os.path.join(HERE, name + ".pt")
        torch.save(fx, path)
        print(name, "->", path, f"{os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
