"""Data-parallel host logic on CPU: world_size 2, gloo backend (SURVEY section 8e semantics)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(nn.Module):
    """Two heads behind a shared trunk: only the routed head receives a gradient (like the 27 task heads)."""

    def __init__(self):
        super().__init__()
        self.trunk = nn.Linear(8, 8)
        self.heads = nn.ModuleDict({"a": nn.Linear(8, 2), "b": nn.Linear(8, 3)})
        self.task_id_to_name = {"a": "Regression", "b": "Regression"}

    def forward(self, x, task_id):
        return self.heads[task_id](torch.tanh(self.trunk(x)))


class _FakeCore:
    """Stands in for SwinCore: a flat gradient buffer reduced stage by stage from the backward hook."""

    def __init__(self, n):
        self._stage_slices = [(0, 8), (8, 24), (24, 40), (40, 56), (56, n)]
        self._p = [nn.Parameter(torch.zeros(n))]

    def ordered_params(self):
        return self._p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mtus_b200 as m
    from mtus_b200 import encoders as enc_mod
    torch.manual_seed(0)
    model = _Toy()
    # ---- 1. GradAllReducer.finish: mean of per-rank grads; params without grad stay None --------------------
    x = torch.randn(4, 8, generator=torch.Generator().manual_seed(10 + rank))
    y = torch.randn(4, 2, generator=torch.Generator().manual_seed(20 + rank))
    loss = (model(x, "a") - y).square().mean()
    loss.backward()
    local = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    red = m.GradAllReducer(model)
    with red:
        pass
    red.finish()
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    ok = True
    for k, p in model.named_parameters():
        if k.startswith("heads.b"):
            ok &= p.grad is None                     # idle head: never reduced into existence
            continue
        mean = sum(g[k] for g in gathered) / world
        ok &= torch.allclose(p.grad, mean, atol=1e-7)
    # ---- 2. trainer step keeps replicas identical and skips the idle head in AdamW ---------------------------
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.1)
    before_b = model.heads["b"].weight.clone()
    tr = m.DataParallelTrainer(model, opt, {"Regression": nn.MSELoss()}, gradient_clip=1.0)
    for step in range(3):
        xs = torch.randn(4, 8, generator=torch.Generator().manual_seed(100 + 7 * step + rank))
        ys = torch.randn(4, 2, generator=torch.Generator().manual_seed(200 + 7 * step + rank))
        tr.step(xs, ys, "a")
    ok &= torch.equal(model.heads["b"].weight, before_b)      # no grad -> no decay, no moment update
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    both = [None] * world
    dist.all_gather_object(both, flat)
    ok &= all(torch.equal(both[0], b) for b in both)
    # ---- 3. stage hook: chunks reduced as backward finishes them, averaged once at the end ----------------------
    n = 64
    core = _FakeCore(n)
    holder = nn.Module()
    holder.encoder = nn.Module()
    holder.encoder.model = core
    red2 = m.GradAllReducer(holder)
    g = torch.full((n,), float(rank + 1))
    with red2:
        hook = enc_mod._STAGE_GRAD_HOOK
        ok &= hook is not None
        for lo_stage in (3, 2, 1, 0):
            s_lo, s_hi = core._stage_slices[lo_stage + 1]
            if lo_stage == 0:
                s_lo = core._stage_slices[0][0]
            hook(g, s_lo, s_hi)
    ok &= enc_mod._STAGE_GRAD_HOOK is None
    ok &= torch.allclose(g, torch.full((n,), sum(range(1, world + 1)) / world))
    # ---- 4. active head reduced in place through one flat gradient block (prepare), buffers averaged ----------
    torch.manual_seed(1 + rank)                          # replicas start DIFFERENT: the trainer must broadcast rank 0
    model2 = _Toy()
    model2.heads["a"] = nn.Sequential(nn.Linear(8, 2), nn.BatchNorm1d(2))
    opt2 = torch.optim.SGD(model2.parameters(), lr=0.1)
    tr2 = m.DataParallelTrainer(model2, opt2, {"Regression": nn.MSELoss()}, gradient_clip=0.0)
    p0 = [None] * world
    dist.all_gather_object(p0, torch.cat([p.detach().flatten() for p in model2.parameters()]))
    ok &= torch.equal(p0[0], p0[1])                      # broadcast at construction
    xs = torch.randn(4, 8, generator=torch.Generator().manual_seed(300 + rank))
    ys = torch.randn(4, 2, generator=torch.Generator().manual_seed(400 + rank))
    ref = _Toy()
    ref.heads["a"] = nn.Sequential(nn.Linear(8, 2), nn.BatchNorm1d(2))
    ref.load_state_dict(model2.state_dict())
    (ref(xs, "a") - ys).square().mean().backward()
    local2 = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    g2 = [None] * world
    dist.all_gather_object(g2, local2)
    tr2.optimizer = type("NoStep", (), {"zero_grad": lambda self_: opt2.zero_grad(), "step": lambda self_: None})()
    tr2.step(xs, ys, "a")
    for k, p in model2.named_parameters():
        if k.startswith("heads.b"):
            ok &= p.grad is None
            continue
        ok &= torch.allclose(p.grad, sum(g[k] for g in g2) / world, atol=1e-7)
    blk = tr2.reducer._head_blocks[id(model2.heads["a"])]
    ok &= all(p.grad.data_ptr() == v.data_ptr() for p, v in blk[1])    # gradients ARE views of the flat block
    rm = [None] * world
    dist.all_gather_object(rm, model2.heads["a"][1].running_mean.clone())
    ok &= torch.equal(rm[0], rm[1])                      # BatchNorm statistics stay identical across replicas
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sampler_needs_a_seed_when_distributed():
    import pytest
    import mtus_b200 as m
    with pytest.raises(ValueError):
        m.DistributedTaskSampler(["a", "b"] * 8, 2, rank=0, world_size=2, seed=None)
    a = list(m.DistributedTaskSampler(["a", "b"] * 8, 2, rank=0, world_size=2, seed=5))
    b = list(m.DistributedTaskSampler(["a", "b"] * 8, 2, rank=1, world_size=2, seed=5))
    assert len(a) == len(b) and all(set(x).isdisjoint(y) for x, y in zip(a, b))


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


# ---- validation split (evaluate(), code/metrics/__init__.py:72-184): batches strided over ranks, metrics gathered ----------
def _fake_evaluate(task_of_sample, batches, batch_ids):
    """The bookkeeping of the reference's evaluate(): per batch, group samples by task id and append ONE value per
    (task, metric) -- here a deterministic function of the sample indices, so any reordering or loss shows."""
    metrics, ids = {}, {}
    for g, batch in zip(batch_ids, batches):
        for tid in sorted({task_of_sample[i] for i in batch}):
            mine = [i for i in batch if task_of_sample[i] == tid]
            mm = metrics.setdefault(tid, {})
            mm.setdefault("Score", []).append(sum((i * 37) % 11 for i in mine) / len(mine))
            if tid == "cls":
                mm.setdefault("F1-Score", []).append(float(len(mine)))
            ids.setdefault(tid, []).append(g)
    return metrics, ids


def _eval_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mtus_b200 as m
    n, B = 53, 4                                          # ragged: 14 batches, the last one short; rank 1 gets no "rare" sample
    task_of_sample = ["seg" if i % 3 else "cls" for i in range(n)]
    task_of_sample[8] = "rare"
    single = m.DistributedEvalBatchSampler(n, B, 0, 1)
    want, _ = _fake_evaluate(task_of_sample, list(single), single.global_batch_ids())
    sampler = m.DistributedEvalBatchSampler(n, B, rank, world)
    got, ids = _fake_evaluate(task_of_sample, list(sampler), sampler.global_batch_ids())
    merged = m.gather_task_metrics(got, ids)
    ok = (merged == want) if rank == 0 else (merged is None)
    covered = [None] * world
    dist.all_gather_object(covered, [i for b in sampler for i in b])
    ok &= sorted(sum(covered, [])) == list(range(n))       # every sample exactly once, no padding
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_validation_split_reproduces_the_single_process_metric_lists():
    import pytest
    import mtus_b200 as m
    s = m.DistributedEvalBatchSampler(10, 4, 0, 1)
    assert list(s) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]] and len(s) == 3
    assert [len(m.DistributedEvalBatchSampler(10, 4, r, 2)) for r in (0, 1)] == [2, 1]
    assert list(m.DistributedEvalBatchSampler(0, 4, 0, 2)) == []
    with pytest.raises(ValueError):
        m.DistributedEvalBatchSampler(10, 4, 2, 2)
    with pytest.raises(ValueError):
        m.gather_task_metrics({"a": {"Dice": [0.5, 0.6]}}, [0])          # one value per batch id
    assert m.gather_task_metrics({"a": {"Dice": [0.5, 0.6]}}, [3, 1]) == {"a": {"Dice": [0.6, 0.5]}}    # world 1: sorted by batch
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_eval_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
