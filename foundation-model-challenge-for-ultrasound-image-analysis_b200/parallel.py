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


class GradAllReducer:
    """Mean all-reduce of the gradients of ``model`` over ``group`` (NCCL on GPUs, gloo in CPU tests)."""
    _avg = None                                 # True: ReduceOp.AVG (NCCL); False: SUM then scale (gloo)

    def __init__(self, model: torch.nn.Module, group=None, overlap_encoder: bool = True):
        self.model, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap_encoder
        self._handles = []
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

    def _avg_op(self):
        # NCCL averages in the collective; gloo (CPU tests) has no AVG: sum, then scale
        if self._avg is None:
            backend = dist.get_backend(self.group) if dist.is_initialized() else "gloo"
            self._avg = backend == "nccl"
        return dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM

    # called from SwinCore._run_backward after each chunk of blocks: flat_grad[lo:hi] is final
    def _stage_hook(self, flat_grad: torch.Tensor, lo: int, hi: int):
        if self.world > 1:
            op = self._avg_op()
            self._handles.append(dist.all_reduce(flat_grad[lo:hi], op=op, group=self.group, async_op=True))
            if lo == self._first_lo:            # last chunk: the gradient views are handed to autograd next
                for h in self._handles:
                    h.wait()                    # stream-ordered wait, not a host sync (NCCL)
                self._handles = []
                if not self._avg:
                    flat_grad.mul_(1.0 / self.world)
        self._enc_reduced = True

    def __enter__(self):
        self._enc_reduced = False
        if self.overlap and self._enc_params and self.world > 1:
            core = self.model.encoder.model
            self._first_lo = core._stage_slices[0][0]
            _enc._STAGE_GRAD_HOOK = self._stage_hook
        return self

    def __exit__(self, *exc):
        _enc._STAGE_GRAD_HOOK = None
        return False

    def finish(self):
        """All-reduce what backward did not already reduce (FPN, heads; the encoder too when not overlapped).

        Flat modules (FPN decoders, the encoder) are reduced in place through their one contiguous gradient block;
        the remaining (head) gradients are coalesced into one call."""
        if self.world == 1:
            return
        op = self._avg_op()
        done = set(self._enc_params) if self._enc_reduced else set()
        handles, scale_later = [], []
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
            handles.append(dist.all_reduce(g, op=op, group=self.group, async_op=True))
            scale_later.append(g)
            done |= {id(p) for p in ps}
        grads = [p.grad for p in self._all_params if p.grad is not None and id(p) not in done]
        flat = None
        if grads:
            flat = torch.cat([g.reshape(-1) for g in grads])
            handles.append(dist.all_reduce(flat, op=op, group=self.group, async_op=True))
        for h in handles:
            h.wait()
        if not self._avg:
            for g in scale_later:
                g.mul_(1.0 / self.world)
            if flat is not None:
                flat.mul_(1.0 / self.world)
        if flat is not None:
            torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split_with_sizes([g.numel() for g in grads]), grads)])


class DataParallelTrainer:
    """One training step with the reference's semantics (train.py:326,440-455) on 1..N ranks."""

    def __init__(self, model, optimizer, loss_functions, loss_weights=None, gradient_clip: float = 1.0, group=None):
        self.model, self.optimizer = model, optimizer
        self.loss_functions, self.loss_weights = loss_functions, loss_weights or {}
        self.clip = float(gradient_clip)
        self.reducer = GradAllReducer(model, group=group)

    def step(self, images, labels, task_id):
        from .losses import compute_task_loss
        task_name = self.model.task_id_to_name[task_id]
        outputs = self.model(images, task_id=task_id)
        loss = compute_task_loss(self.loss_functions, task_name, outputs, labels)
        total = loss * self.loss_weights.get(task_name, 1.0)
        self.optimizer.zero_grad()
        with self.reducer:
            total.backward()
        self.reducer.finish()
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
