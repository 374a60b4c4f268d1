"""Pins the oracle (CPU only): Swin restatement vs torchvision's independent implementation, FPN restatement vs a
functional re-derivation, committed golden tensors (generated through the reference's own model code,
tests/golden/make_golden.py), structural invariants, and -- where /root/reference exists -- the
reference's MultiTaskModel running on the shims."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, HAS_REFERENCE

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


# ---- Swin vs torchvision -------------------------------------------------------------------------
@pytest.mark.parametrize("name,img", [("swin_tiny_patch4_window7_224", 224), ("swin_micro_patch4_window7_test", 224)])
def test_swin_oracle_matches_torchvision(name, img):
    from torchvision.models.swin_transformer import SwinTransformer
    from oracle import swin
    torch.manual_seed(0)
    o = swin.create_model(name, features_only=True, img_size=img, drop_path_rate=0.0).eval()
    ed, depths, heads, win = swin.SWIN_VARIANTS[name]
    tv = SwinTransformer(patch_size=[4, 4], embed_dim=ed, depths=list(depths), num_heads=list(heads),
                         window_size=[win, win], stochastic_depth_prob=0.0).features.eval()
    missing = tv.load_state_dict(swin.to_torchvision_state_dict(o.state_dict()), strict=False)
    assert not [k for k in missing.missing_keys if "relative_position_index" not in k], missing
    assert not missing.unexpected_keys
    x = torch.randn(2, 3, img, img, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        fo = o(x)
        h, ft = x, []
        for i, layer in enumerate(tv):
            h = layer(h)
            if i in (1, 3, 5, 7):
                ft.append(h)
    for a, b in zip(fo, ft):
        assert a.shape == b.shape
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), float((a - b).abs().max())


def test_swin_structural_invariants():
    from oracle import swin
    for w in (7, 12):
        idx = swin.relative_position_index(w, w)
        n = w * w
        assert idx.shape == (n, n)
        assert (idx.diagonal() == ((2 * w - 1) ** 2 - 1) // 2).all()
        assert idx.min() == 0 and idx.max() == (2 * w - 1) ** 2 - 1
    blk = swin.SwinTransformerBlock(32, (14, 14), 1, 7, 3)
    m = blk.get_attn_mask()
    assert m.shape == (4, 49, 49) and set(m.unique().tolist()) == {-100.0, 0.0}
    # window clipped + shift disabled when the resolution is not larger than the window (stage 4 at 224)
    blk = swin.SwinTransformerBlock(32, (7, 7), 1, 7, 3)
    assert blk.get_attn_mask() is None and tuple(blk.shift_size) == (0, 0)
    blk = swin.SwinTransformerBlock(32, (4, 4), 1, 7, 3)
    assert tuple(blk.window_size) == (4, 4)


def test_swin_roll_then_pad_order():
    """timm rolls THEN zero-pads; pad-then-roll (torchvision / HF) gives a different result."""
    from oracle import swin
    torch.manual_seed(0)
    blk = swin.SwinTransformerBlock(32, (9, 9), 1, 7, 3).eval()
    x = torch.randn(1, 9, 9, 32)
    y = blk._attn(blk.norm1(x))
    # manual restatement of the order
    s = 3
    h = torch.roll(blk.norm1(x), (-s, -s), (1, 2))
    h = F.pad(h, (0, 0, 0, 5, 0, 5))
    wins = swin.window_partition(h, (7, 7)).view(-1, 49, 32)
    a = blk.attn(wins, mask=blk.get_attn_mask())
    h = swin.window_reverse(a.view(-1, 7, 7, 32), (7, 7), 14, 14)[:, :9, :9]
    ref = torch.roll(h, (s, s), (1, 2))
    assert torch.allclose(y, ref, atol=1e-6)


# ---- FPN vs functional re-derivation ---------------------------------------------------------------
def test_fpn_oracle_matches_functional_rederivation():
    from oracle import fpn
    torch.manual_seed(0)
    dec = fpn.FPNDecoder([3, 32, 64, 128, 256], 4, 64, 32, 0.0, "cat").eval()
    feats = [torch.randn(2, c, s, s) for c, s in ((32, 16), (64, 8), (128, 4), (256, 2))]
    sd = dec.state_dict()
    c2, c3, c4, c5 = feats

    def conv(x, k, pad=0):
        return F.conv2d(x, sd[k + ".weight"], sd.get(k + ".bias"), padding=pad)

    p5 = conv(c5, "p5")
    p4 = F.interpolate(p5, scale_factor=2.0, mode="nearest") + conv(c4, "p4.skip_conv")
    p3 = F.interpolate(p4, scale_factor=2.0, mode="nearest") + conv(c3, "p3.skip_conv")
    p2 = F.interpolate(p3, scale_factor=2.0, mode="nearest") + conv(c2, "p2.skip_conv")
    outs = []
    for i, (p, nup) in enumerate(((p5, 3), (p4, 2), (p3, 1), (p2, 0))):
        h = p
        for l in range(max(nup, 1)):
            pre = f"seg_blocks.{i}.block.{l}.block"
            h = F.relu(F.group_norm(conv(h, pre + ".0", 1), 32, sd[pre + ".1.weight"], sd[pre + ".1.bias"], 1e-5))
            if nup > 0:
                h = F.interpolate(h, scale_factor=2.0, mode="bilinear", align_corners=True)
        outs.append(h)
    ref = torch.cat(outs, dim=1)
    with torch.no_grad():
        y = dec(feats)
    assert y.shape == (2, 128, 16, 16) and dec.out_channels == 128
    assert torch.allclose(y, ref, rtol=1e-5, atol=1e-6)
    dec_add = fpn.FPNDecoder([3, 32, 64, 128, 256], 4, 64, 32, 0.0, "add").eval()
    assert dec_add.out_channels == 32
    with pytest.raises(ValueError):
        fpn.FPNDecoder([3, 32, 64, 128, 256], 4, 64, 32, 0.0, "mul")


# ---- goldens -----------------------------------------------------------------------------------------
def _golden_case(name):
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden
    fx = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    cfg, oracle, x = make_golden.build_case(name, fx["spec"])
    for k, v in oracle.state_dict().items():
        if v.is_floating_point():
            s, a = fx["weight_checksums"][k]
            assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, a), f"weight RNG drift in {k}"
    return fx, cfg, oracle, x, make_golden.subsample


@pytest.mark.parametrize("name", ["micro_64_4types", "config1_swin_t_224"])
def test_oracle_reproduces_reference_goldens(name):
    fx, cfg, oracle, x, sub = _golden_case(name)
    spec = fx["spec"]
    with torch.no_grad():
        feats = oracle.encoder(x)
    for f, g, (s, a) in zip(feats, fx["features"], fx["feature_sums"]):
        assert torch.allclose(sub(f, spec["sub"]), g, rtol=1e-5, atol=1e-6)
        assert abs(float(f.double().sum()) - s) <= 1e-6 * a
    for tid in spec["tasks"]:
        oracle.zero_grad(set_to_none=True)
        out = oracle(x, tid)
        assert torch.allclose(sub(out.detach(), spec["sub"]), fx["outputs"][tid], rtol=1e-5, atol=1e-6)
        out.float().square().mean().backward()
        norms = fx["grad_norms"][tid]
        got = {k: p.grad for k, p in oracle.named_parameters() if p.grad is not None}
        assert set(got) == set(norms)
        for k, g in got.items():
            assert abs(float(g.double().norm()) - norms[k]) <= 1e-4 * max(norms[k], 1e-12) + 1e-12, k
        for k, g in fx["grads"][tid].items():
            assert torch.allclose(got[k], g, rtol=1e-4, atol=1e-9), k


# ---- the reference's own classes on the shims (authoring container only) -------------------------------
@pytest.mark.skipif(not HAS_REFERENCE, reason="/root/reference is not present on this box")
def test_reference_multitask_model_on_shims_equals_oracle_model():
    import mtus_b200 as m
    from oracle import shims
    from oracle.model import OracleMultiTaskModel
    models, _, _ = shims.import_reference_models("/root/reference")
    tasks = [t for t in m.tasks_27() if t["task_id"] in ("T2A_fetal_head", "T2C_fetal_head", "T3A_breast_tumor",
                                                         "T4A_fetal_femur", "T5_fetal_abdomen")]
    cfg = m.make_config("swin_micro_patch4_window7_test", 64, 2, tasks=tasks, dropout=0.0, mixed_precision=False)
    torch.manual_seed(0)
    oracle = OracleMultiTaskModel(cfg, drop_path_rate=0.0).eval()
    ref = models.build_model(cfg).eval()
    ref.load_state_dict(oracle.state_dict(), strict=True)
    assert list(ref.state_dict().keys()) == list(oracle.state_dict().keys())
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    for t in tasks:
        with torch.no_grad():
            a, b = ref(x, t["task_id"]), oracle(x, t["task_id"])
        assert torch.equal(a, b), t["task_id"]
    with pytest.raises(ValueError):
        ref(x, "no_such_task")
    with pytest.raises(ValueError):
        oracle(x, "no_such_task")


@pytest.mark.skipif(not HAS_REFERENCE, reason="needs /root/reference (authoring container)")
@pytest.mark.parametrize("embedding", [False, True])
def test_film_reference_model_equals_oracle_and_native_key_layout(embedding):
    """model.use_film (best_config.yaml:58): the reference's own MultiTaskModel + film_layer.py on the shims == the oracle's
    FiLM restatement bit for bit, and the native model exposes the same film_generator.* state-dict keys."""
    import mtus_b200 as m
    from oracle import shims
    from oracle.model import OracleMultiTaskModel
    models, _, _ = shims.import_reference_models("/root/reference")
    tasks = [t for t in m.tasks_27() if t["task_id"] in ("T2C_fetal_head", "T3A_breast_tumor", "T4A_fetal_femur")]
    cfg = m.make_config("swin_micro_patch4_window7_test", 64, 2, tasks=tasks, dropout=0.0, mixed_precision=False)
    cfg.config["model"]["use_film"] = True
    cfg.config["model"]["film"] = {"use_task_embedding": embedding, "embedding_dim": 16, "use_affine": True}
    torch.manual_seed(0)
    oracle = OracleMultiTaskModel(cfg, drop_path_rate=0.0).eval()
    with torch.no_grad():
        for k, p in oracle.named_parameters():
            if k.startswith("film_generator."):
                p.add_(torch.randn_like(p) * 0.3)        # gamma = 1, beta = 0 at init would hide a missing modulation
    ref = models.build_model(cfg).eval()
    ref.load_state_dict(oracle.state_dict(), strict=True)
    assert list(ref.state_dict().keys()) == list(oracle.state_dict().keys())
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    for t in tasks:
        with torch.no_grad():
            assert torch.equal(ref(x, t["task_id"]), oracle(x, t["task_id"])), t["task_id"]
    native = m.build_model(cfg, precision="fp32")
    assert list(native.state_dict().keys()) == list(oracle.state_dict().keys())
    native.load_state_dict(oracle.state_dict(), strict=True)
    enc, head = native.get_trainable_parameters()
    ids = {id(p) for p in enc + head}
    assert not any(id(p) in ids for p in native.film_generator.parameters())     # as in the reference (:282-308)


def _task_prompt_cfg(m, tasks, image, mode, channels, names=None, mixed=False):
    cfg = m.make_config("swin_micro_patch4_window7_test", image, 2, tasks=tasks, dropout=0.0, mixed_precision=mixed)
    cfg.set_task_configs_from_dataset(tasks)                 # the reference refuses TaskPrompt2D on YAML-only task lists
    cfg.config["model"]["task_prompt"] = {"enabled": True, "channels": channels, "prompt_size": 8, "inject_mode": mode,
                                          "init_scale": 0.1, "use_tanh": True, "apply_to_task_names": names}
    return cfg


@pytest.mark.skipif(not HAS_REFERENCE, reason="needs /root/reference (authoring container)")
@pytest.mark.parametrize("mode,channels,names", [("add", 1, None), ("mul", 3, ["segmentation"])])
def test_task_prompt_reference_model_equals_oracle_and_native_key_layout(mode, channels, names):
    """model.task_prompt (multitask_model.py:81-111, 194-199; task_prompt.py): the reference's own MultiTaskModel +
    TaskPrompt2D on the shims == the oracle's restatement bit for bit (forward and the prompt's gradients), the product's
    TaskPrompt2D == the reference's class, and the native model has the same task_prompt.* keys and parameter groups."""
    import mtus_b200 as m
    from mtus_b200.task_prompt import TaskPrompt2D, build_task_prompt_metadata
    from oracle import shims
    from oracle.model import OracleMultiTaskModel
    models, _, _ = shims.import_reference_models("/root/reference")
    tasks = [t for t in m.tasks_27() if t["task_id"] in ("T2C_fetal_head", "T3A_breast_tumor", "T4A_fetal_femur")]
    cfg = _task_prompt_cfg(m, tasks, 64, mode, channels, names)
    torch.manual_seed(0)
    oracle = OracleMultiTaskModel(cfg, drop_path_rate=0.0).eval()
    with torch.no_grad():
        oracle.task_prompt.prompt_scale.fill_(0.7)
    ref = models.build_model(cfg).eval()
    ref.load_state_dict(oracle.state_dict(), strict=True)
    assert list(ref.state_dict().keys()) == list(oracle.state_dict().keys())
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    for t in tasks:
        ref.zero_grad(set_to_none=True)
        oracle.zero_grad(set_to_none=True)
        yr, yo = ref(x, t["task_id"]), oracle(x, t["task_id"])
        assert torch.equal(yr, yo), t["task_id"]
        yr.square().mean().backward()
        yo.square().mean().backward()
        applies = names is None or t["task_name"].lower() in names
        for (k, pr), (_, po) in zip(ref.task_prompt.named_parameters(), oracle.task_prompt.named_parameters()):
            assert (pr.grad is not None) == applies, (k, t["task_id"])
            if applies:
                assert torch.equal(pr.grad, po.grad), k
    # the product's module against the reference's class (same seed -> same parameters -> same map, 1 ulp of the Linear)
    ref_tp = models.TaskPrompt2D
    meta_ref = sys.modules[ref_tp.__module__].build_task_prompt_metadata(m.tasks_27())
    meta = build_task_prompt_metadata(m.tasks_27())
    assert torch.equal(meta[0], meta_ref[0]) and meta[1] == meta_ref[1] and meta[2] == meta_ref[2]
    torch.manual_seed(1)
    a = TaskPrompt2D(m.tasks_27(), channels, 8, mode, 0.3, True)
    torch.manual_seed(1)
    b = ref_tp(m.tasks_27(), channels, 8, mode, 0.3, True)
    assert all(torch.equal(a.state_dict()[k], v) for k, v in b.state_dict().items())
    for tid in ("T1_fetal_planes", "T2A_fetal_abdomen"):
        torch.testing.assert_close(a.apply(x, tid), b.apply(x, tid), rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        a.apply(x, "no_such_task")
    native = m.build_model(cfg, precision="fp32")
    assert list(native.state_dict().keys()) == list(oracle.state_dict().keys())
    native.load_state_dict(oracle.state_dict(), strict=True)
    enc, head = native.get_trainable_parameters()
    enc_r, head_r = ref.get_trainable_parameters()
    assert len(enc) == len(enc_r) and len(head) == len(head_r)
    assert all(any(p is q for q in head) for p in native.task_prompt.parameters())          # multitask_model.py:305-306


# ---- second independent pin of the Swin restatement: Hugging Face transformers -------------------
def _hf_state_dict_from_oracle(sd, depths):
    """oracle / timm keys -> transformers.SwinModel keys (qkv split into query / key / value; HF stores the
    PatchMerging of stage i+1 at the END of its layer i)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("patch_embed.proj."):
            out["embeddings.patch_embeddings.projection." + k.rsplit(".", 1)[1]] = v
        elif k.startswith("patch_embed.norm."):
            out["embeddings.norm." + k.rsplit(".", 1)[1]] = v
        elif k.startswith("layers_"):
            i = int(k[len("layers_")])
            rest = k.split(".", 1)[1]
            if rest.startswith("downsample."):
                out[f"encoder.layers.{i - 1}.{rest}"] = v
                continue
            _, j, tail = rest.split(".", 2)
            pre = f"encoder.layers.{i}.blocks.{j}."
            leaf = tail.rsplit(".", 1)[1] if "." in tail else tail
            if tail.startswith("norm1."):
                out[pre + "layernorm_before." + leaf] = v
            elif tail.startswith("norm2."):
                out[pre + "layernorm_after." + leaf] = v
            elif tail == "attn.relative_position_bias_table":
                out[pre + "attention.self.relative_position_bias_table"] = v
            elif tail.startswith("attn.qkv."):
                c = v.shape[0] // 3
                for n, name in enumerate(("query", "key", "value")):
                    out[pre + f"attention.self.{name}.{leaf}"] = v[n * c:(n + 1) * c]
            elif tail.startswith("attn.proj."):
                out[pre + "attention.output.dense." + leaf] = v
            elif tail.startswith("mlp.fc1."):
                out[pre + "intermediate.dense." + leaf] = v
            elif tail.startswith("mlp.fc2."):
                out[pre + "output.dense." + leaf] = v
    return out


@pytest.mark.parametrize("img", [224, 448])
def test_swin_oracle_matches_huggingface_transformers(img):
    """Unpadded configurations only: HF pads then rolls, timm rolls then pads.  224 -> maps 56/28/14/7 (last
    stage unshifted, one window); 448 -> 112/56/28/14 (shifted windows and masks in all four stages)."""
    transformers = pytest.importorskip("transformers")
    from transformers import SwinConfig, SwinModel
    from oracle import swin
    name = "swin_micro_patch4_window7_test"
    ed, depths, heads, win = swin.SWIN_VARIANTS[name]
    torch.manual_seed(0)
    o = swin.create_model(name, features_only=True, img_size=img, drop_path_rate=0.0).eval()
    with torch.no_grad():
        for p in o.parameters():
            p.add_(torch.randn_like(p) * 0.02)          # biases / LayerNorm affine away from their (0, 1) init
    cfg = SwinConfig(image_size=img, patch_size=4, num_channels=3, embed_dim=ed, depths=list(depths), num_heads=list(heads),
                     window_size=win, mlp_ratio=4.0, qkv_bias=True, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0, drop_path_rate=0.0, hidden_act="gelu",
                     use_absolute_embeddings=False, layer_norm_eps=1e-5)
    hf = SwinModel(cfg, add_pooling_layer=False).eval()
    res = hf.load_state_dict(_hf_state_dict_from_oracle(o.state_dict(), depths), strict=False)
    assert not res.unexpected_keys
    assert not [k for k in res.missing_keys if "relative_position_index" not in k and not k.startswith("layernorm.")], res
    x = torch.randn(1 if img > 224 else 2, 3, img, img, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        fo = o(x)
        emb, dims = hf.embeddings(x)
        enc = hf.encoder(emb, dims, output_hidden_states=True, output_hidden_states_before_downsampling=True, return_dict=True)
    hs = enc.reshaped_hidden_states[1:]                  # [B, C, H, W] per stage, before the next PatchMerging
    assert len(hs) == 4
    for a, b in zip(fo, hs):
        b = b.permute(0, 2, 3, 1)
        assert a.shape == b.shape
        assert torch.allclose(a, b, rtol=1e-4, atol=2e-5), float((a - b).abs().max())


def test_swin_oracle_padded_windows_match_huggingface_on_unshifted_blocks():
    """Window PADDING pinned where the two orderings coincide: without a cyclic shift, pad-then-roll (HF, torchvision) and
    roll-then-pad (timm) are the same computation, so a model whose blocks are all unshifted (one block per stage) must agree
    with HF on maps that are not multiples of the window -- 256 px -> 64 / 32 / 16 / 8, padded to 70 / 35 / 21 / 14 (zero rows
    after the norm, so padded tokens carry the qkv bias and ARE attended).  What stays unpinned against
    an independent implementation is only the order of roll and pad inside SHIFTED blocks of padded maps
    (test_swin_roll_then_pad_order pins that against a hand-built case)."""
    transformers = pytest.importorskip("transformers")
    from transformers import SwinConfig, SwinModel
    from oracle import swin
    name = "swin_unshifted_1_1_1_1_test"
    ed, depths, heads, win = swin.SWIN_VARIANTS.setdefault(name, (32, (1, 1, 1, 1), (1, 2, 4, 8), 7))
    img = 256
    torch.manual_seed(0)
    o = swin.create_model(name, features_only=True, img_size=img, drop_path_rate=0.0).eval()
    with torch.no_grad():
        for p in o.parameters():
            p.add_(torch.randn_like(p) * 0.05)          # non-zero qkv biases: the padded tokens' keys / values matter
    cfg = SwinConfig(image_size=img, patch_size=4, num_channels=3, embed_dim=ed, depths=list(depths), num_heads=list(heads),
                     window_size=win, mlp_ratio=4.0, qkv_bias=True, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0, drop_path_rate=0.0, hidden_act="gelu",
                     use_absolute_embeddings=False, layer_norm_eps=1e-5)
    hf = SwinModel(cfg, add_pooling_layer=False).eval()
    res = hf.load_state_dict(_hf_state_dict_from_oracle(o.state_dict(), depths), strict=False)
    assert not res.unexpected_keys
    assert not [k for k in res.missing_keys if "relative_position_index" not in k and not k.startswith("layernorm.")], res
    x = torch.randn(2, 3, img, img, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        fo = o(x)
        emb, dims = hf.embeddings(x)
        enc = hf.encoder(emb, dims, output_hidden_states=True, output_hidden_states_before_downsampling=True, return_dict=True)
    hs = enc.reshaped_hidden_states[1:]
    assert [tuple(f.shape[1:3]) for f in fo] == [(64, 64), (32, 32), (16, 16), (8, 8)]
    for a, b in zip(fo, hs):
        b = b.permute(0, 2, 3, 1)
        assert a.shape == b.shape
        assert torch.allclose(a, b, rtol=1e-4, atol=2e-5), float((a - b).abs().max())


# ---- FPN lateral + top-down pathway vs torchvision.ops.FeaturePyramidNetwork ------------------------------------
def test_fpn_lateral_topdown_matches_torchvision_feature_pyramid_network():
    """smp's p5 = 1x1(c5), p_k = nearest_up(p_{k+1}) + 1x1(c_k) is torchvision's FPN inner pathway.  torchvision then
    applies a 3x3 `layer_block` per level (smp has none): with identity kernels there its outputs ARE p2..p5, which the
    oracle must reproduce from the same 1x1 weights."""
    from collections import OrderedDict
    from torchvision.ops import FeaturePyramidNetwork
    from oracle.fpn import FPNDecoder
    chans = [32, 64, 128, 256]
    torch.manual_seed(0)
    dec = FPNDecoder(encoder_channels=[3] + chans, encoder_depth=4, pyramid_channels=64, segmentation_channels=32,
                     dropout=0.0, merge_policy="cat").eval()
    tv = FeaturePyramidNetwork(chans, 64).eval()
    with torch.no_grad():
        lat = [dec.p2.skip_conv, dec.p3.skip_conv, dec.p4.skip_conv, dec.p5]
        for blk, conv in zip(tv.inner_blocks, lat):
            blk[0].weight.copy_(conv.weight)
            blk[0].bias.copy_(conv.bias)
        for blk in tv.layer_blocks:
            blk[0].weight.zero_()
            blk[0].bias.zero_()
            for c in range(64):
                blk[0].weight[c, c, 1, 1] = 1.0
    g = torch.Generator().manual_seed(1)
    feats = [torch.randn(2, c, s, s, generator=g) for c, s in zip(chans, (32, 16, 8, 4))]
    with torch.no_grad():
        out = tv(OrderedDict((str(i), f) for i, f in enumerate(feats)))
        p5 = dec.p5(feats[3])
        p4 = dec.p4(p5, feats[2])
        p3 = dec.p3(p4, feats[1])
        p2 = dec.p2(p3, feats[0])
    for a, b in zip((p2, p3, p4, p5), out.values()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5), float((a - b).abs().max())
