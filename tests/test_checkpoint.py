"""Checkpoint / pretrained-weight interchange, CPU only: key maps from timm (current and pre-0.9
layouts), from torchvision, the reference's ``encoder.model.`` prefix, bicubic relative-position-table resize, the
``pretrained: imagenet`` local-file lookup and the ``best_model.pth`` round trip."""
import os

import pytest
import torch
import torch.nn.functional as F


def _core(name="swin_micro_patch4_window7_test", img=224):
    import mtus_b200 as m
    torch.manual_seed(3)
    return m.SwinCore(name, img_size=img, precision="fp32")


def _timm_full_model_sd(core):
    """A timm SwinTransformer (classifier) state dict carrying the native module's values: layers.{i}. keys plus the
    tensors a real checkpoint also holds (final norm, head, index / mask buffers)."""
    sd = {k.replace("layers_", "layers.", 1) if k.startswith("layers_") else k: v.clone() for k, v in core.state_dict().items()}
    c_last = core.embed_dim * 8
    sd["norm.weight"], sd["norm.bias"] = torch.ones(c_last), torch.zeros(c_last)
    sd["head.fc.weight"], sd["head.fc.bias"] = torch.zeros(1000, c_last), torch.zeros(1000)
    sd["layers.0.blocks.0.attn.relative_position_index"] = torch.zeros(49, 49, dtype=torch.long)
    sd["layers.0.blocks.1.attn_mask"] = torch.zeros(64, 49, 49)
    return sd


def test_timm_layouts_and_prefixes_map_onto_the_native_keys():
    from mtus_b200 import checkpoint as ck
    src = _core()
    want = {k: v.clone() for k, v in src.state_dict().items()}
    full = _timm_full_model_sd(src)
    # current timm layout
    dst = _core()
    dst.load_state_dict(ck.convert_swin_state_dict(full, dst))
    for k, v in dst.state_dict().items():
        assert torch.equal(v, want[k]), k
    # pre-0.9 layout: PatchMerging stored at the END of stage i-1
    old = {}
    for k, v in full.items():
        if ".downsample." in k:
            i = int(k.split(".")[1])
            k = k.replace(f"layers.{i}.", f"layers.{i - 1}.", 1)
        old[k] = v
    assert any(k.startswith("layers.0.downsample.") for k in old)
    dst = _core()
    dst.load_state_dict(ck.convert_swin_state_dict(old, dst))
    for k, v in dst.state_dict().items():
        assert torch.equal(v, want[k]), k
    # the reference's own checkpoint prefix (MultiTaskModel.state_dict(): encoder.model.layers_0...)
    ref = {"encoder.model." + k: v for k, v in src.state_dict().items()}
    dst = _core()
    dst.load_state_dict(ck.convert_swin_state_dict(ref, dst))
    for k, v in dst.state_dict().items():
        assert torch.equal(v, want[k]), k
    # a tensor missing from the file is an error, not a silent random init
    bad = dict(full)
    bad.pop("layers.2.blocks.1.mlp.fc1.weight")
    with pytest.raises(KeyError):
        ck.convert_swin_state_dict(bad, dst)


def test_torchvision_weights_load_and_reproduce_torchvision_features():
    """torchvision's ImageNet Swin files are the other common source of `imagenet` weights: its state dict maps onto
    the native keys and the ORACLE loaded with the converted weights reproduces torchvision's own features."""
    from torchvision.models.swin_transformer import SwinTransformer
    from mtus_b200 import checkpoint as ck
    from oracle import swin
    name = "swin_micro_patch4_window7_test"
    ed, depths, heads, win = swin.SWIN_VARIANTS[name]
    torch.manual_seed(0)
    tv = SwinTransformer(patch_size=[4, 4], embed_dim=ed, depths=list(depths), num_heads=list(heads),
                         window_size=[win, win], stochastic_depth_prob=0.0).eval()
    for p in tv.parameters():                      # torchvision zero-inits some biases: make every tensor informative
        torch.nn.init.normal_(p, std=0.05)
    core = _core(name)
    conv = ck.convert_swin_state_dict(tv.state_dict(), core)
    core.load_state_dict(conv)
    o = swin.create_model(name, features_only=True, img_size=224, drop_path_rate=0.0).eval()
    o.load_state_dict(core.state_dict())
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        fo = o(x)
        h, ft = x, []
        for i, layer in enumerate(tv.features):
            h = layer(h)
            if i in (1, 3, 5, 7):
                ft.append(h)
    for a, b in zip(fo, ft):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), float((a - b).abs().max())


def test_relative_position_table_is_resized_bicubically_when_the_window_differs():
    from mtus_b200 import checkpoint as ck
    t7 = torch.randn(169, 4, generator=torch.Generator().manual_seed(0))
    t12 = ck.resize_rel_pos_bias_table(t7, 12)
    assert t12.shape == (529, 4)
    ref = F.interpolate(t7.t().reshape(1, 4, 13, 13), size=(23, 23), mode="bicubic", align_corners=False).reshape(4, 529).t()
    assert torch.allclose(t12, ref)
    assert torch.equal(ck.resize_rel_pos_bias_table(t7, 7), t7)
    # end to end: window-7 weights into a model whose last stage clips the window (img 128 -> 4x4 map, 7x7 table)
    src = _core(img=224)
    dst = _core(img=128)
    sd = ck.convert_swin_state_dict(_timm_full_model_sd(src), dst)
    k = "layers_3.blocks.0.attn.relative_position_bias_table"
    assert sd[k].shape == (49, 8) and src.state_dict()[k].shape == (169, 8)
    dst.load_state_dict(sd)


def test_pretrained_imagenet_uses_a_local_file_and_fails_loudly_without_one(tmp_path, monkeypatch):
    import mtus_b200 as m
    monkeypatch.setenv("MTUS_PRETRAINED_DIR", str(tmp_path))
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("HOME", str(tmp_path))
    with pytest.raises(RuntimeError, match="no local file"):
        m.SwinTransformerEncoder("swin_micro_patch4_window7_test", pretrained=True, precision="fp32")
    src = _core()
    torch.save(_timm_full_model_sd(src), os.path.join(tmp_path, "swin_micro_patch4_window7_test.pth"))
    cfg = m.make_config("swin_micro_patch4_window7_test", 224, 2)
    cfg.config["model"]["encoder"]["pretrained"] = "imagenet"        # configs/swin_b.yaml:41
    enc = m.build_encoder(cfg, precision="fp32")
    for k, v in enc.model.state_dict().items():
        assert torch.equal(v, src.state_dict()[k]), k
    # explicit path form
    cfg.config["model"]["encoder"]["pretrained"] = os.path.join(tmp_path, "swin_micro_patch4_window7_test.pth")
    enc2 = m.build_encoder(cfg, precision="fp32")
    assert enc2.pretrained_path.endswith(".pth")


def test_best_model_round_trip_and_periodic_checkpoint_dict(tmp_path):
    """train.py:695 saves model.state_dict(); :714-727 saves a dict with model_state_dict; :735 reloads."""
    import mtus_b200 as m
    from mtus_b200 import checkpoint as ck
    tasks = [t for t in m.tasks_27() if t["task_id"] in ("T2A_fetal_abdomen", "T1_fetal_planes", "T4A_fetal_brain")]
    cfg = m.make_config("swin_micro_patch4_window7_test", 64, 2, tasks=tasks)
    torch.manual_seed(0)
    a = m.build_model(cfg, precision="fp32")
    torch.manual_seed(1)
    b = m.build_model(cfg, precision="fp32")
    p1, p2 = os.path.join(tmp_path, "best_model.pth"), os.path.join(tmp_path, "checkpoint_epoch_5.pth")
    torch.save(a.state_dict(), p1)
    torch.save({"epoch": 5, "model_state_dict": a.state_dict(), "best_val_score": 0.5}, p2)
    for path in (p1, p2):
        torch.manual_seed(1)
        b = m.build_model(cfg, precision="fp32")
        ck.load_checkpoint(b, path)
        for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
            assert ka == kb and torch.equal(va, vb), ka
    assert any(k.startswith("encoder.model.layers_0.blocks.0.attn.qkv") for k in a.state_dict())
