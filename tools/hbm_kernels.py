"""Host-free timings of the bandwidth-bound kernels (kbench graphs), one line per kernel: us, GB/s, fraction of the HBM peak.

    python tools/hbm_kernels.py [name-filter]
    MTUS_B200_SO=<package>/libmtus_b200_diag.so python tools/hbm_kernels.py     # -DMTUS_DIAG_NOATOM build: atomic tails removed
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import mtus_b200  # noqa: F401


def main():
    from mtus_b200 import kbench as K
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak = 6548.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    cases = [
        ("ln_fwd s1 [100352,128]", lambda: K.time_layernorm_fwd(100352, 128, dev)),
        ("ln_fwd s3 [6272,512]", lambda: K.time_layernorm_fwd(6272, 512, dev)),
        ("ln_bwd s1 [100352,128]", lambda: K.time_layernorm_bwd(100352, 128, dev)),
        ("ln_bwd s2 [25088,256]", lambda: K.time_layernorm_bwd(25088, 256, dev)),
        ("ln_bwd s3 [6272,512]", lambda: K.time_layernorm_bwd(6272, 512, dev)),
        ("ln_bwd s4 [1568,1024]", lambda: K.time_layernorm_bwd(1568, 1024, dev)),
        ("merge_ln_fwd [32,56,56,128]", lambda: K.time_patch_merge_ln(32, 56, 128, dev)),
        ("merge_ln_bwd [32,56,56,128]", lambda: K.time_patch_merge_ln(32, 56, 128, dev, backward=True)),
        ("merge_ln_bwd [32,28,28,256]", lambda: K.time_patch_merge_ln(32, 28, 256, dev, backward=True)),
        ("merge_ln_bwd [32,14,14,512]", lambda: K.time_patch_merge_ln(32, 14, 512, dev, backward=True)),
        ("merge_ln_fwd [32,14,14,512]", lambda: K.time_patch_merge_ln(32, 14, 512, dev)),
        ("colsum [100352,256]", lambda: K.time_colsum(100352, 256, dev)),
        ("colsum [25088,256]", lambda: K.time_colsum(25088, 256, dev)),
        ("gn_fwd fused [32,56,56,128]", lambda: K.time_groupnorm_relu(32, 56, 128, dev, fused=True)),
        ("gn_bwd relu [32,56,56,128]", lambda: K.time_groupnorm_bwd(32, 56, 128, dev, act=0)),
        ("gn_bwd silu [32,56,56,128]", lambda: K.time_groupnorm_bwd(32, 56, 128, dev, act=1)),
        ("gn_bwd relu [32,28,28,128]", lambda: K.time_groupnorm_bwd(32, 28, 128, dev, act=0)),
        ("bilinear [32,28,28,128]", lambda: K.time_bilinear(32, 28, 128, dev)),
        ("merge_fwd [32,56,56,4x128]", lambda: K.time_fpn_merge(32, 56, 128, 4, dev)),
        ("attn_fwd s3", lambda: K.time_window_attn(32, 14, 512, 16, 7, 3, dev)[:2]),
        ("attn_bwd s3", lambda: K.time_window_attn(32, 14, 512, 16, 7, 3, dev, backward=True)[:2]),
        ("attn_bwd s1", lambda: K.time_window_attn(32, 56, 128, 4, 7, 3, dev, backward=True)[:2]),
    ]
    print("library:", os.environ.get("MTUS_B200_SO", "libmtus_b200.so"))
    for name, fn in cases:
        if flt and flt not in name:
            continue
        t, by = fn()
        print(f"{name:34s} {t * 1e6:8.2f} us  {by / t / 1e9:8.1f} GB/s  {by / t / 1e9 / peak:5.3f}")


if __name__ == "__main__":
    main()
