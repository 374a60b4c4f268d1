"""Host-side mirror of the reference interface (CPU only): factories, attributes, state-dict keys, error
behaviour, sampler semantics.  No kernel is launched here."""
import random

import pytest
import torch

import mtus_b200 as m


def _cfg(**kw):
    tasks = [t for t in m.tasks_27() if t["task_id"] in ("T2A_fetal_abdomen", "T1_fetal_planes", "T4A_fetal_brain",
                                                         "T5_fetal_femur")]
    return m.make_config("swin_micro_patch4_window7_test", 64, 2, tasks=tasks, **kw)


def test_config_interface():
    cfg = m.swin_b_27task()
    assert cfg.get("model.encoder.name") == "swin_b"
    assert cfg.get("model.encoder.pretrained") is None
    assert cfg.get("no.such.key", 7) == 7
    tasks = cfg.get_task_configs()
    assert len(tasks) == 27
    kinds = [t["task_name"] for t in tasks]
    assert (kinds.count("segmentation"), kinds.count("classification"), kinds.count("detection"),
            kinds.count("Regression")) == (12, 9, 3, 3)
    assert cfg.get_loss_config("detection")["type"] == "Detection"


def test_encoder_factory_contract():
    """encoders.py:665-691 / :37-160: name map, attributes, channels, strict pretrained handling."""
    assert m.SWIN_MODEL_MAPPING["swin_b"] == "swin_base_patch4_window7_224"
    cfg = m.make_config("swin_t", 224, 4)
    enc = m.build_encoder(cfg)
    assert enc.is_timm_encoder and enc.output_stride == 32
    assert enc.out_channels == [3, 96, 192, 384, 768]
    assert enc.supports_task_id is False and enc.handles_moe is False and enc.use_moe is False
    assert enc.get_moe_stats() == []
    assert enc.model.feature_info.channels() == [96, 192, 384, 768]
    cfg.config["model"]["encoder"]["pretrained"] = "imagenet"
    with pytest.raises(RuntimeError):
        m.build_encoder(cfg)
    cfg.config["model"]["encoder"]["pretrained"] = None
    cfg.config["model"]["encoder"]["name"] = "vit_b"
    with pytest.raises(NotImplementedError):
        m.build_encoder(cfg)
    with pytest.raises(RuntimeError):
        m.SwinTransformerEncoder("swin_nonexistent", pretrained=False)


def test_decoder_factory_contract():
    """decoders.py:63-103: keys, aliasing when not separate, out_channels for cat / add."""
    cfg = _cfg(separate_fpn=False)
    enc = m.build_encoder(cfg)
    dec = m.build_decoders(enc, cfg)
    assert set(dec) == {"fpn_seg", "fpn_det", "fpn_cls", "fpn_reg"}
    assert dec["fpn_det"] is dec["fpn_seg"] and dec["fpn_cls"] is dec["fpn_seg"] and dec["fpn_reg"] is dec["fpn_seg"]
    assert dec["fpn_seg"].out_channels == 512
    dec = m.build_decoders(enc, _cfg(separate_fpn=True, merge_policy="add"))
    assert dec["fpn_det"] is not dec["fpn_seg"] and dec["fpn_seg"].out_channels == 128
    with pytest.raises(ValueError):
        m.FPNDecoder([3, 32, 64, 128, 256], 4, merge_policy="mul")


def test_state_dict_keys_match_the_reference_layout():
    """Checkpoints are bare state_dicts (train.py:695): keys and shapes must equal the oracle's (= timm / smp names)."""
    from oracle.model import OracleMultiTaskModel
    cfg = _cfg(dropout=0.0)
    torch.manual_seed(0)
    oracle = OracleMultiTaskModel(cfg)
    model = m.build_model(cfg, precision="fp32")
    so, sm = oracle.state_dict(), model.state_dict()
    # buffers that are non-persistent in timm (relative_position_index, attn_mask) must not appear
    assert set(sm) == set(so), sorted(set(sm) ^ set(so))[:10]
    for k in so:
        assert tuple(so[k].shape) == tuple(sm[k].shape), k
    model.load_state_dict(so, strict=True)
    flat = model.encoder.model.flat_params()
    w = dict(model.named_parameters())["encoder.model.layers_1.blocks.0.attn.qkv.weight"]
    assert w.data_ptr() >= flat.data_ptr() and torch.equal(w, so["encoder.model.layers_1.blocks.0.attn.qkv.weight"])
    assert "encoder.model.layers_0.downsample.norm.weight" not in sm          # merge sits at the START of stages 1..3
    assert "encoder.model.layers_1.downsample.reduction.weight" in sm
    assert "fpn_decoder_seg.p4.skip_conv.weight" in sm and "fpn_decoder_seg.seg_blocks.0.block.2.block.1.bias" in sm


def test_routing_and_errors_without_a_gpu():
    cfg = _cfg()
    model = m.build_model(cfg, precision="fp32")
    x = torch.zeros(2, 3, 64, 64)
    with pytest.raises(ValueError, match="Unknown task_id"):
        model(x, "nope")                                                     # multitask_model.py:187-188
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(x, "T1_fetal_planes")                                          # the product path never runs on the CPU
    enc_p, head_p = model.get_trainable_parameters()
    assert len(enc_p) == len(list(model.encoder.parameters()))
    ids = {id(p) for p in enc_p} | {id(p) for p in head_p}
    # decoders that no task routes through (cls / reg when use_fpn_for_* is false) are left out, as in the
    # reference (multitask_model.py:294-303)
    unused = {id(p) for d in (model.fpn_decoder_cls, model.fpn_decoder_reg) for p in d.parameters()}
    assert ids == {id(p) for p in model.parameters()} - unused
    model.freeze_encoder()
    assert not any(p.requires_grad for p in model.encoder.parameters())
    model.unfreeze_encoder()
    assert all(p.requires_grad for p in model.encoder.parameters())
    # FiLM is provided (fused into the merge kernel) and so is TaskPrompt2D (test_task_prompt_wiring_without_a_gpu); MoE is
    # outside the hot path and must fail loudly
    c2 = _cfg()
    c2.config["model"]["use_film"] = True
    film_model = m.build_model(c2, precision="fp32")
    assert any(k.startswith("film_generator.task_gammas.") for k in film_model.state_dict())
    c3 = _cfg()
    c3.config["model"]["moe"] = {"enabled": True}
    with pytest.raises(NotImplementedError):
        m.build_model(c3)


def test_sampler_is_task_synchronous_and_partitions_the_global_batch():
    """MultiTaskUniformSampler (data/dataset.py:140-192) made distributed: same task on every rank each step,
    disjoint rank slices, rank-identical wrap-around reshuffle."""
    rng = random.Random(0)
    task_of = [rng.choice(["a", "b", "c"]) for _ in range(500)]
    world, bs = 4, 8
    samplers = [m.DistributedTaskSampler(task_of, bs, rank=r, world_size=world, steps_per_epoch=40, seed=42)
                for r in range(world)]
    assert len(samplers[0]) == 40
    for batches in zip(*[iter(s) for s in samplers]):
        tasks = {task_of[i] for b in batches for i in b}
        assert len(tasks) == 1
        flat = [i for b in batches for i in b]
        assert all(len(b) == bs for b in batches)
        # a global batch only repeats an index when a task is exhausted mid-batch
        assert len(set(flat)) >= len(flat) - bs
    s1 = m.DistributedTaskSampler(task_of, bs, 0, 1, steps_per_epoch=10, seed=7)
    s2 = m.DistributedTaskSampler(task_of, bs, 0, 1, steps_per_epoch=10, seed=7)
    assert list(s1) == list(s2)
    with pytest.raises(ValueError):
        m.DistributedTaskSampler(task_of, bs, rank=4, world_size=4)


def test_synthetic_batches_have_the_documented_shapes():
    g = torch.Generator().manual_seed(0)
    for t in m.tasks_27()[::5]:
        x, y = m.synthetic_batch(t, 3, 64, generator=g)
        assert x.shape == (3, 3, 64, 64)
        if t["task_name"] == "segmentation":
            assert y.shape == (3, 64, 64) and y.dtype == torch.int64
        elif t["task_name"] == "classification":
            assert y.shape == (3,) and int(y.max()) < t["num_classes"]
        elif t["task_name"] == "detection":
            assert y.shape == (3, 4) and (y[:, 2] > y[:, 0]).all() and (y[:, 3] > y[:, 1]).all()
        else:
            assert y.shape == (3, 2 * t["num_classes"])


def test_backward_chunk_plan_tiles_the_flat_gradient():
    """Data-parallel chunk plan (SwinCore._backward_chunks): block ranges descend from the top, never cross a stage,
    and their gradient slices tile the flat gradient block exactly once (so every chunk can be all-reduced as soon as
    its backward range has run)."""
    enc = m.SwinTransformerEncoder("swin_b", pretrained=False, img_size=224, precision="fp32")
    core = enc.model
    for k in (1, 3, 5, 100):
        chunks = core._backward_chunks(k)
        nblk = sum(core.depths)
        assert chunks[0][0] == nblk and chunks[-1][1] == 0
        for (hi, lo, _, _), (hi2, _, _, _) in zip(chunks, chunks[1:]):
            assert lo == hi2 and hi > lo
        bounds, acc = [], 0
        for d in core.depths:
            bounds.append((acc, acc + d))
            acc += d
        for hi, lo, _, _ in chunks:
            assert any(b0 <= lo and hi <= b1 for b0, b1 in bounds), "a chunk crosses a stage"
            assert hi - lo <= k
        spans = sorted((g_lo, g_hi) for _, _, g_lo, g_hi in chunks)
        assert spans[0][0] == core._stage_slices[0][0]
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0, "gradient slices must be contiguous and disjoint"
        assert spans[-1][1] == core._stage_slices[4][1]
        # every parameter lies inside exactly one slice, together with the rest of its block
        for name, (p, off, numel, shape) in core._params_by_name.items():
            inside = [s for s in spans if s[0] <= off and off + numel <= s[1]]
            assert len(inside) == 1, name


def test_persistent_buffers_keep_autograd_semantics():
    """Buffers the executors' CUDA-graph cache keys on (FlatParamModule): the gradient block is reused only when no
    parameter still holds a gradient (so gradient accumulation never sees its views overwritten), and training
    workspaces cycle through a pool keyed by (batch, size, device)."""
    enc = m.SwinTransformerEncoder("swin_t", pretrained=False, img_size=224, precision="fp32")
    core = enc.model
    dev = torch.device("cpu")
    g1 = core._grad_block(dev)
    assert g1.numel() == core._n_flat and float(g1.abs().sum()) == 0.0
    g1.add_(1.0)
    g2 = core._grad_block(dev)                      # no parameter holds a gradient: same block, zeroed again
    assert g2.data_ptr() == g1.data_ptr() and float(g2.abs().sum()) == 0.0
    p0 = core.ordered_params()[0]
    p0.grad = g2[:p0.numel()].view_as(p0)           # a backward handed out views: accumulation must not clobber them
    g3 = core._grad_block(dev)
    assert g3.data_ptr() != g2.data_ptr()
    p0.grad = None
    assert core._grad_block(dev).data_ptr() == g2.data_ptr()
    # workspace pool
    w1 = core._take_workspace(4, 1024, dev)
    w2 = core._take_workspace(4, 1024, dev)         # two forwards before a backward: two distinct workspaces
    assert w1.data_ptr() != w2.data_ptr()
    core._return_workspace(4, w1)
    assert core._take_workspace(4, 1024, dev).data_ptr() == w1.data_ptr()
    assert core._take_workspace(8, 1024, dev).data_ptr() != w1.data_ptr()


def test_parameter_version_key_sees_every_in_place_write_through_torch():
    """FlatParamModule.params_version_key guards the optimizer-written bf16 shadow (encoders.py): it must change when a
    parameter view or the flat block is written in place (load_state_dict, torch optimizers, broadcasts) and stay put
    otherwise; a re-flatten (module move) changes it through the block's address."""
    enc = m.SwinTransformerEncoder("swin_t", pretrained=False, img_size=224, precision="bf16")
    core = enc.model
    k0 = core.params_version_key()
    assert core.params_version_key() == k0                       # reading it does not change it
    with torch.no_grad():
        core.ordered_params()[5].mul_(1.0)
    k1 = core.params_version_key()
    assert k1 != k0
    with torch.no_grad():
        core.flat_params().add_(0.0)
    k2 = core.params_version_key()
    assert k2 != k1
    core.load_state_dict(core.state_dict())
    k3 = core.params_version_key()
    assert k3 != k2
    core.to(torch.float32)                                       # _apply re-flattens: new block
    assert core.params_version_key()[0] != k3[0] or core.params_version_key() != k3


def test_trainer_loss_item_on_cpu_matches_the_returned_loss():
    """DataParallelTrainer.loss_item(): the side-stream read-back degenerates to a plain copy on CPU tensors."""
    import types
    from mtus_b200.parallel import DataParallelTrainer
    tr = DataParallelTrainer.__new__(DataParallelTrainer)
    tr._loss_reader = None
    with pytest.raises(RuntimeError):
        tr.loss_item()
    tr._post_loss_readback(torch.tensor(1.25))
    assert tr.loss_item() == 1.25
    tr._post_loss_readback(torch.tensor(-3.5, dtype=torch.float64))
    assert tr.loss_item() == -3.5


def test_flat_adamw_is_a_torch_optimizer_and_follows_lr_schedulers():
    """the reference's CosineAnnealingLR (code/train.py:222-253, stepped at :698-704) must be able to drive
    FlatAdamW.  CPU part: a model without flat blocks (everything goes through the mirrored torch groups)."""
    import torch
    import torch.nn as nn
    import mtus_b200 as m

    class Toy(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = nn.Linear(6, 6)
            self.heads = nn.Linear(6, 2)

        def get_trainable_parameters(self):
            return list(self.encoder.parameters()), list(self.heads.parameters())

    torch.manual_seed(0)
    a, b = Toy(), Toy()
    b.load_state_dict(a.state_dict())
    oa = m.FlatAdamW(a, lr=1e-2, weight_decay=0.1, max_grad_norm=0.0)
    ob = torch.optim.AdamW([{"params": list(b.encoder.parameters()), "lr": 1e-3},
                            {"params": list(b.heads.parameters()), "lr": 1e-2}], lr=1e-2, weight_decay=0.1)
    assert isinstance(oa, torch.optim.Optimizer) and [g["lr"] for g in oa.param_groups] == [1e-3, 1e-2]
    sa = torch.optim.lr_scheduler.CosineAnnealingLR(oa, T_max=5, eta_min=1e-6)
    sb = torch.optim.lr_scheduler.CosineAnnealingLR(ob, T_max=5, eta_min=1e-6)
    x = torch.randn(8, 6)
    for _ in range(5):
        for mod, opt, sch in ((a, oa, sa), (b, ob, sb)):
            opt.zero_grad()
            mod.heads(torch.tanh(mod.encoder(x))).square().mean().backward()
            opt.step()
            sch.step()
        assert [g["lr"] for g in oa.param_groups] == [g["lr"] for g in ob.param_groups]
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, atol=1e-7)
    sd = oa.state_dict()
    oa.load_state_dict(sd)
    assert [g["lr"] for g in oa.param_groups] == [g["lr"] for g in ob.param_groups]


def test_bench_arms_print_the_same_config_and_non_zero_ranks_of_the_reference_arm_do_no_work(monkeypatch):
    """bench.py contract: the CPU reference arm reports on OUR arm's config / metric / unit (the driver compares the two
    lines), the seeded task sequence is the sampler's, and under torchrun only rank 0 of the reference arm works."""
    import argparse
    import bench
    wl = bench.WORKLOADS["swin_b_224"]
    for world in (1, 2, 8):
        c = bench._config_dict(wl, world, wl["batch"], wl["image"])
        assert c["global_batch"] == 32 * world and c["parallelism"] == f"dp{world}" and c["workload"] == bench.WORKLOAD
        assert "model" not in c
    ids = [t["task_id"] for t in m.tasks_27()]
    rng = random.Random(42)
    assert bench._task_sequence(ids, 16) == [rng.choice(ids) for _ in range(16)]
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    assert bench.run_reference(argparse.Namespace(gpus=2, steps=1, warmup=0)) == 0      # returns before building anything


def test_task_prompt_wiring_without_a_gpu():
    """model.task_prompt (multitask_model.py:81-111): requires dataset-derived task configs, registers task_prompt.* with
    the reference's names, joins the head parameter group (:305-306), refuses the raw uint8 batch, and keeps
    nn.Module.apply(fn) working next to the reference's apply(x, task_id)."""
    tasks = [t for t in m.tasks_27() if t["task_id"] in ("T2C_fetal_head", "T1_fetal_planes")]
    cfg = m.make_config("swin_micro_patch4_window7_test", 64, 2, tasks=tasks, mixed_precision=False)
    cfg.config["model"]["task_prompt"] = {"enabled": True, "channels": 1, "prompt_size": 8, "apply_to_task_names": ["Segmentation"]}
    with pytest.raises(ValueError, match="dataset-derived"):
        m.build_model(cfg, precision="fp32")
    cfg.set_task_configs_from_dataset(tasks)
    model = m.build_model(cfg, precision="fp32")
    assert model.use_task_prompt and model.task_prompt_apply_task_names == {"segmentation"}
    keys = [k for k in model.state_dict() if k.startswith("task_prompt.")]
    assert keys == ["task_prompt.prompt_scale", "task_prompt.task_metadata", "task_prompt.prompt_proj.weight", "task_prompt.prompt_proj.bias"]
    assert model.task_prompt.prompt_proj.weight.shape == (64, model.task_prompt.prompt_dim)
    _, head = model.get_trainable_parameters()
    assert all(any(p is q for q in head) for p in model.task_prompt.parameters())
    with pytest.raises(TypeError, match="uint8"):
        model(torch.zeros(2, 64, 64, 3, dtype=torch.uint8), "T2C_fetal_head")
    seen = []
    model.apply(lambda mod: seen.append(type(mod).__name__))
    assert "TaskPrompt2D" in seen
    x = torch.randn(2, 3, 64, 64)
    y = model.task_prompt.apply(x, "T2C_fetal_head")
    assert y.shape == x.shape and not torch.equal(y, x)
