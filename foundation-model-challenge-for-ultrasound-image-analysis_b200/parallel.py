"""Batch data-parallel training for the hot path: one process per GPU, NCCL over NVLink 5 / NVSwitch.

The reference has NO multi-GPU code; this is the data-parallel wrapper north_star asks for,
preserving the single-process step semantics of ``/root/reference/code/train.py:440-455``:

* ``DistributedTaskSampler`` -- ``MultiTaskUniformSampler`` (code/data/dataset.py:140-192) made
  task-synchronous: every rank draws the SAME task each step from an identically seeded
  ``random.Random`` and takes its own slice of a ``world*batch`` global batch, so the set of
  gradient-bearing parameters is identical on all ranks.
* ``GradAllReducer`` -- mean all-reduce of gradients.  The encoder's flat gradient buffer is reduced
  stage by stage (layers_3 first) on NCCL's stream while the remaining stages still run backward
  (``encoders._STAGE_GRAD_HOOK``); FPN / head gradients are coalesced into one call.  Parameters whose
  grad is ``None`` (the 26 idle heads, idle decoders) are left untouched so AdamW keeps skipping them
  (train.py:440,455 semantics).
* ``DataParallelTrainer`` -- zero_grad -> forward -> loss -> backward -> all-reduce -> clip -> step.
* ``DistributedEvalBatchSampler`` + ``gather_task_metrics`` -- the validation split of ``evaluate()``: batches strided over
  ranks, per-batch metric values gathered to rank 0 in single-process order.
"""

import random
from typing import Dict, Iterator, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import encoders as _enc


class DistributedTaskSampler(torch.utils.data.Sampler):
    """Batch sampler: same task on every rank each step, disjoint per-rank slices of the global batch."""

    def __init__(self, task_ids: Sequence[str], batch_size: int, rank: int = 0, world_size: int = 1,
                 steps_per_epoch: Optional[int] = None, seed: Optional[int] = None):
        self.batch_size, self.rank, self.world_size = int(batch_size), int(rank), int(world_size)
        if not (0 <= self.rank < self.world_size):
            raise ValueError("rank must be in [0, world_size)")
        if self.world_size > 1 and seed is None:
            # random.Random(None) seeds from OS entropy per process: ranks would draw different tasks, hold different
            # sets of gradient-bearing parameters and issue mismatched collectives (hang or wrong sums)
            raise ValueError("DistributedTaskSampler: world_size > 1 needs an explicit seed shared by all ranks")
        self.rng = random.Random(seed)          # identical stream on every rank
        self.indices_by_task: Dict[str, List[int]] = {}
        for idx, tid in enumerate(task_ids):
            self.indices_by_task.setdefault(tid, []).append(idx)
        self.task_ids = list(self.indices_by_task.keys())
        for tid in self.task_ids:
            self.rng.shuffle(self.indices_by_task[tid])
        n = len(task_ids)
        self.steps_per_epoch = steps_per_epoch if steps_per_epoch is not None else n // (self.batch_size * self.world_size)

    def __iter__(self) -> Iterator[List[int]]:
        cursors = {tid: 0 for tid in self.task_ids}
        gb = self.batch_size * self.world_size
        for _ in range(self.steps_per_epoch):
            tid = self.rng.choice(self.task_ids)
            idx = self.indices_by_task[tid]
            start, end = cursors[tid], cursors[tid] + gb
            if end > len(idx):                  # wrap around with a rank-identical reshuffle
                batch = list(idx[start:])
                self.rng.shuffle(idx)
                rem, cur = gb - len(batch), 0
                while rem > 0:                  # a task smaller than one global batch repeats its samples
                    take = idx[:rem]
                    batch.extend(take)
                    rem -= len(take)
                    cur = len(take)
                cursors[tid] = cur
            else:
                batch = idx[start:end]
                cursors[tid] = end
            yield batch[self.rank * self.batch_size:(self.rank + 1) * self.batch_size]

    def __len__(self) -> int:
        return self.steps_per_epoch


class DistributedEvalBatchSampler(torch.utils.data.Sampler):
    """Validation split for ``evaluate()`` (code/metrics/__init__.py:72-184) under data parallelism.

    The reference's ``val_loader`` is a plain ``shuffle=False`` loader (code/train.py:164-171) and every metric is the mean of
    PER-BATCH values (metrics/__init__.py:176-182), so the split is strided over BATCHES, not samples: global batch ``g`` =
    indices ``[g * B, (g + 1) * B)`` goes to rank ``g % world``.  No padding and no repeated samples -- the union of the ranks'
    batches is exactly the single-process batch list, and ``gather_task_metrics`` puts the per-batch values back in that order,
    so rank 0 reports the numbers a single process would.  Use as ``DataLoader(val_dataset, batch_sampler=...)``."""

    def __init__(self, dataset_len: int, batch_size: int, rank: int = 0, world_size: int = 1):
        self.n, self.batch_size, self.rank, self.world_size = int(dataset_len), int(batch_size), int(rank), int(world_size)
        if not (0 <= self.rank < self.world_size) or self.batch_size <= 0:
            raise ValueError("rank must be in [0, world_size) and batch_size positive")
        self.num_global_batches = (self.n + self.batch_size - 1) // self.batch_size

    def global_batch_ids(self) -> List[int]:
        return list(range(self.rank, self.num_global_batches, self.world_size))

    def __iter__(self) -> Iterator[List[int]]:
        for g in self.global_batch_ids():
            yield list(range(g * self.batch_size, min((g + 1) * self.batch_size, self.n)))

    def __len__(self) -> int:
        return len(self.global_batch_ids())


def gather_task_metrics(task_metrics: Dict[str, Dict[str, list]], batch_ids: Sequence[Sequence[int]], group=None, dst: int = 0):
    """Merge the per-rank ``task_metrics[task_id][metric] -> [value per batch]`` dictionaries of ``evaluate()`` on rank ``dst``.

    ``batch_ids[task_id]`` (or one list for all tasks) holds, for each appended value, the GLOBAL index of the batch it came from
    (``DistributedEvalBatchSampler.global_batch_ids()``; a task absent from a batch appends nothing).  Returns, on ``dst``, the
    merged dictionary with every list ordered by global batch index -- element for element the list a single process builds --
    and ``None`` on the other ranks.  Host-side Python objects over the group's object collective (gloo or NCCL)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    per_task_ids = batch_ids if isinstance(batch_ids, dict) else {t: list(batch_ids) for t in task_metrics}
    local = {t: {"ids": list(per_task_ids[t]), "metrics": {k: list(v) for k, v in mm.items()}} for t, mm in task_metrics.items()}
    for t, e in local.items():
        for k, v in e["metrics"].items():
            if len(v) != len(e["ids"]):
                raise ValueError(f"gather_task_metrics: {t}/{k} has {len(v)} values for {len(e['ids'])} batch ids")
    if world == 1:
        parts = [local]
    else:
        parts = [None] * world if rank == dst else None
        dist.gather_object(local, parts, dst=dst, group=group)
        if rank != dst:
            return None
    merged: Dict[str, Dict[str, list]] = {}
    for t in sorted({t for part in parts for t in part}):
        names = []
        for part in parts:
            names += [k for k in part.get(t, {"metrics": {}})["metrics"] if k not in names]
        merged[t] = {}
        for k in names:
            rows = [(g, v) for part in parts if t in part and k in part[t]["metrics"]
                    for g, v in zip(part[t]["ids"], part[t]["metrics"][k])]
            merged[t][k] = [v for _, v in sorted(rows, key=lambda r: r[0])]
    return merged


class GradAllReducer:
    """Mean all-reduce of the gradients of ``model`` over ``group`` (NCCL on GPUs, gloo in CPU tests).

    Everything is reduced IN PLACE through contiguous gradient blocks, asynchronously on the collective's own stream,
    while the encoder (the last and longest part of backward) still runs:

    * heads: ``prepare(modules)`` points the ``.grad`` of the active head's parameters at views of one persistent flat
      block before backward (autograd accumulates into them), so the head's gradient is one buffer -- no concatenate /
      copy-back round trip;
    * FPN decoders own a flat gradient block already (``FlatParamModule``);
    * both are complete when the encoder's backward begins (``encoders._PRE_BACKWARD_HOOK``) and are reduced there;
    * the encoder's flat block is reduced chunk by chunk as its backward proceeds (``encoders._STAGE_GRAD_HOOK``).

    ``finish()`` reduces whatever the hooks did not see (frozen / absent encoder, foreign parameters) and makes the
    current stream wait for all outstanding collectives."""
    _avg = None                                 # True: ReduceOp.AVG (NCCL); False: SUM then scale (gloo)

    def __init__(self, model: torch.nn.Module, group=None, overlap_encoder: bool = True):
        self.model, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap_encoder
        self._handles = []
        self._scale_later = []
        self._done = set()
        self._enc_params = set()
        enc = getattr(model, "encoder", None)
        core = getattr(enc, "model", None)
        if core is not None and hasattr(core, "ordered_params"):
            self._enc_params = {id(p) for p in core.ordered_params()}
        self._enc_reduced = False
        self._avg = None
        from ._native import FlatParamModule
        self._flat_modules = [m for m in model.modules() if isinstance(m, FlatParamModule)]
        self._all_params = list(model.parameters())
        self._head_blocks = {}                  # id(module) -> (flat buffer, [(param, view)])
        self._prepared = []

    def _avg_op(self):
        # NCCL averages in the collective; gloo (CPU tests) has no AVG: sum, then scale
        if self._avg is None:
            backend = dist.get_backend(self.group) if dist.is_initialized() else "gloo"
            self._avg = backend == "nccl"
        return dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM

    def _reduce_async(self, buf: torch.Tensor):
        self._handles.append(dist.all_reduce(buf, op=self._avg_op(), group=self.group, async_op=True))
        if not self._avg:
            self._scale_later.append(buf)

    def _wait_all(self):
        for h in self._handles:
            h.wait()                            # stream-ordered wait, not a host sync (NCCL)
        self._handles = []
        for b in self._scale_later:
            b.mul_(1.0 / self.world)
        self._scale_later = []

    # ---- head gradients as one flat block --------------------------------------------------------------------
    def prepare(self, modules):
        """Call after zero_grad and before backward with the non-flat modules that will receive gradients this step."""
        self._prepared = []
        if self.world == 1:
            return
        for mod in modules:
            ent = self._head_blocks.get(id(mod))
            ps = [p for p in mod.parameters() if p.requires_grad]
            if not ps:
                continue
            if ent is None or ent[0].device != ps[0].device or len(ent[1]) != len(ps):
                sizes = [((p.numel() + 3) // 4) * 4 for p in ps]
                flat = torch.zeros(sum(sizes), dtype=ps[0].dtype, device=ps[0].device)
                views, off = [], 0
                for p, n in zip(ps, sizes):
                    # same sizes AND strides as the parameter (channels-last conv weights): fused optimizers require
                    # gradient and parameter layouts to agree
                    dense = p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)
                    v = flat[off:off + p.numel()]
                    views.append((p, v.as_strided(p.shape, p.stride()) if dense else v.view_as(p)))
                    off += n
                ent = self._head_blocks[id(mod)] = (flat, views)
            flat, views = ent
            if any(p.dtype != flat.dtype for p, _ in views):
                continue                        # mixed dtypes: leave this module to the generic path in finish()
            flat.zero_()
            for p, v in views:
                p.grad = v
            self._prepared.append(ent)

    # called from SwinCore._run_backward before its first chunk: heads and decoders have finished their backward
    def _pre_backward_hook(self):
        if self.world > 1:
            self._reduce_early()

    def _reduce_early(self):
        for flat, views in self._prepared:
            if all(p.grad is v or (p.grad is not None and p.grad.data_ptr() == v.data_ptr()) for p, v in views):
                self._reduce_async(flat)
                self._done |= {id(p) for p, _ in views}
        self._prepared = []
        for mod in self._flat_modules:
            g = getattr(mod, "_last_flat_grad", None)
            ps = mod.ordered_params()
            if g is None or not ps or id(ps[0]) in self._done or id(ps[0]) in self._enc_params:
                continue
            self._reduce_async(g)
            self._done |= {id(p) for p in ps}

    # called from SwinCore._run_backward after each chunk of blocks: flat_grad[lo:hi] is final
    def _stage_hook(self, flat_grad: torch.Tensor, lo: int, hi: int):
        if self.world > 1:
            self._reduce_async(flat_grad[lo:hi])
            if lo == self._first_lo:            # last chunk: the gradient views are handed to autograd next
                self._wait_all()
        self._enc_reduced = True

    def __enter__(self):
        self._enc_reduced = False
        self._done = set()
        if self.overlap and self._enc_params and self.world > 1:
            core = self.model.encoder.model
            self._first_lo = core._stage_slices[0][0]
            _enc._STAGE_GRAD_HOOK = self._stage_hook
            _enc._PRE_BACKWARD_HOOK = self._pre_backward_hook
        return self

    def __exit__(self, *exc):
        _enc._STAGE_GRAD_HOOK = None
        _enc._PRE_BACKWARD_HOOK = None
        return False

    def finish(self):
        """Reduce what backward did not already reduce and wait (stream-ordered) for every outstanding collective."""
        if self.world == 1:
            return
        done = self._done | (set(self._enc_params) if self._enc_reduced else set())
        self._done = done
        self._reduce_early()                    # prepared heads / flat decoders the pre-backward hook did not see
        done = self._done
        for mod in self._flat_modules:
            g = getattr(mod, "_last_flat_grad", None)
            ps = mod.ordered_params()
            if g is None or not ps or id(ps[0]) in done:
                continue
            first = ps[0].grad
            if first is None:
                continue
            if not (g.data_ptr() <= first.data_ptr() < g.data_ptr() + g.numel() * g.element_size()):
                continue                        # autograd copied the views: fall through to the generic path
            self._reduce_async(g)
            done |= {id(p) for p in ps}
        grads = [p.grad for p in self._all_params if p.grad is not None and id(p) not in done]
        flat = None
        if grads:                               # foreign parameters: one coalesced call and a copy back
            flat = torch.cat([g.reshape(-1) for g in grads])
            self._reduce_async(flat)
        self._wait_all()
        if flat is not None:
            torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split_with_sizes([g.numel() for g in grads]), grads)])

    # ---- replica consistency ---------------------------------------------------------------------------------
    def broadcast_parameters(self, src: int = 0):
        """Parameters and buffers of every rank := those of ``src`` (what DDP does at construction)."""
        if self.world == 1:
            return
        with torch.no_grad():
            seen = set()
            for mod in self._flat_modules:
                dist.broadcast(mod.flat_params(), src=src, group=self.group)
                seen |= {id(p) for p in mod.ordered_params()}
            for p in self._all_params:
                if id(p) not in seen:
                    dist.broadcast(p.data, src=src, group=self.group)
            for b in self.model.buffers():
                dist.broadcast(b.data, src=src, group=self.group)

    def sync_buffers(self, modules=None):
        """Average the floating-point buffers (BatchNorm running statistics of the detection heads) over the ranks and
        take rank 0's integer buffers, so replicas -- and the rank-0 checkpoint -- stay identical."""
        if self.world == 1:
            return
        mods = modules if modules is not None else [self.model]
        with torch.no_grad():
            for mod in mods:
                for b in mod.buffers():
                    if b.is_floating_point():
                        dist.all_reduce(b.data, op=dist.ReduceOp.SUM, group=self.group)
                        b.data.mul_(1.0 / self.world)
                    else:
                        dist.broadcast(b.data, src=0, group=self.group)


class _DoneNow:
    def synchronize(self):
        pass


class DataParallelTrainer:
    """One training step with the reference's semantics (train.py:326,440-455) on 1..N ranks.

    At construction every rank takes rank 0's parameters and buffers; each step the active head's buffers (BatchNorm
    running statistics) are averaged over the ranks.  ``gradient_clip`` is THE clip threshold: an optimizer that clips
    inside its step (``FlatAdamW``) is given this value (0 disables clipping, as in train.py:444-446)."""

    def __init__(self, model, optimizer, loss_functions, loss_weights=None, gradient_clip: float = 1.0, group=None):
        self.model, self.optimizer = model, optimizer
        self.loss_functions, self.loss_weights = loss_functions, loss_weights or {}
        self.clip = float(gradient_clip)
        if getattr(optimizer, "handles_clipping", False):
            optimizer.max_norm = self.clip
        self.reducer = GradAllReducer(model, group=group)
        self.reducer.broadcast_parameters(0)
        self._loss_reader = None

    def loss_item(self) -> float:
        """The last step's loss as a Python float (the reference's ``loss.item()`` per step, train.py:330-332).

        ``tensor.item()`` on the returned loss is ordered behind EVERYTHING the step enqueued (backward, all-reduce,
        optimizer), so the host would sit idle until the step ends and the device would then wait for the next step's first
        launches.  Here the value is copied to pinned host memory on a side stream that only waits for the forward + loss
        (an event recorded before ``backward``): the host gets the number while the backward is still running and enqueues
        the next step behind it."""
        if self._loss_reader is None or self._loss_reader[2] is None:
            raise RuntimeError("loss_item(): no step has run yet")
        host, _, done = self._loss_reader
        done.synchronize()
        return float(host[0])

    def _post_loss_readback(self, total):
        if not total.is_cuda:
            self._loss_reader = (total.detach().reshape(1).float().clone(), None, _DoneNow())
            return
        if self._loss_reader is None or self._loss_reader[1] is None or self._loss_reader[1].device != total.device:
            self._loss_reader = (torch.empty(1, dtype=torch.float32).pin_memory(), torch.cuda.Stream(device=total.device), None)
        host, side, _ = self._loss_reader
        cur = torch.cuda.current_stream(total.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        src = total.detach()
        with torch.cuda.stream(side):
            side.wait_event(ready)
            host.copy_(src.reshape(1).float(), non_blocking=True)
            done = torch.cuda.Event()
            done.record(side)
        src.record_stream(side)
        self._loss_reader = (host, side, done)

    def step(self, images, labels, task_id):
        from .losses import compute_task_loss
        task_name = self.model.task_id_to_name[task_id]
        outputs = self.model(images, task_id=task_id)
        loss = compute_task_loss(self.loss_functions, task_name, outputs, labels)
        total = loss * self.loss_weights.get(task_name, 1.0)
        self._post_loss_readback(total)
        self.optimizer.zero_grad()
        head = self.model.heads[task_id] if hasattr(self.model, "heads") else None
        if head is not None:
            self.reducer.prepare([head])
        with self.reducer:
            total.backward()
        self.reducer.finish()
        if head is not None and self.reducer.world > 1 and any(True for _ in head.buffers()):
            self.reducer.sync_buffers([head])
        if self.clip > 0 and not getattr(self.optimizer, "handles_clipping", False):
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        self.optimizer.step()
        return total.detach()


class DevicePrefetcher:
    """Host -> device copies of the NEXT batch on a side stream while the current step computes.

    The reference moves every batch with a blocking ``images.to(device)`` right before the forward
    (``code/train.py:305``); with pinned host memory the copy of batch i+1 can run under step i instead::

        pf = DevicePrefetcher(device)
        pf.issue(x0, y0)
        for i in range(n):
            x, y = pf.take()                 # current stream waits for the copy (no host sync)
            if i + 1 < n: pf.issue(x_next, y_next)
            trainer.step(x, y, task_id)
    """

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._slot = None

    def issue(self, *host_tensors):
        if self._slot is not None:
            raise RuntimeError("DevicePrefetcher: the previous batch was not taken")
        with torch.cuda.stream(self.stream):
            dev = [t.to(self.device, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._slot = (dev, ev)

    def pending(self) -> bool:
        return self._slot is not None

    def take(self):
        if self._slot is None:
            raise RuntimeError("DevicePrefetcher: nothing was issued")
        dev, ev = self._slot
        self._slot = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)             # allocated on the copy stream, consumed on the compute stream
        return dev


def synthetic_batch(task_cfg, batch, image_size, generator=None, device="cpu", dtype=torch.float32):
    """Synthetic ultrasound-shaped batch for one task: N(0,1) images and labels by task type."""
    name, n = task_cfg["task_name"], task_cfg["num_classes"]
    x = torch.randn(batch, 3, image_size, image_size, generator=generator, dtype=torch.float32).to(dtype)
    if name == "segmentation":
        y = torch.randint(0, n, (batch, image_size, image_size), generator=generator)
    elif name == "classification":
        y = torch.randint(0, n, (batch,), generator=generator)
    elif name == "detection":
        a, b = torch.rand(batch, 2, generator=generator), torch.rand(batch, 2, generator=generator)
        lo, hi = torch.minimum(a, b), (torch.maximum(a, b) + 1e-3).clamp(max=1.0)
        y = torch.stack([lo[:, 0], lo[:, 1], hi[:, 0], hi[:, 1]], dim=1)
    else:
        y = torch.rand(batch, 2 * n, generator=generator)
    return x.to(device), y.to(device)
