"""Shared plumbing for the two native modules (Swin encoder core, FPN decoder).

Both keep their parameters as ``nn.Parameter`` *views* into one flat fp32 buffer whose layout is
dictated by the C++ executors (``mtus_*_param_info``): the state-dict keys stay those of timm / smp
(the reference checkpoints with ``model.state_dict()``, code/train.py:695), while the kernels see one
contiguous parameter block and write one contiguous gradient block (one NCCL call per stage).
"""

import ctypes as C
from typing import List, Tuple

import torch
import torch.nn as nn

from . import _lib


def precision_to_dtype(precision: str):
    if precision in ("bf16", "bfloat16"):
        return _lib.BF16, torch.bfloat16
    if precision in ("fp32", "float32"):
        return _lib.F32, torch.float32
    raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")


def enumerate_params(info_fn, cfg) -> List[Tuple[str, int, Tuple[int, ...]]]:
    out = []
    name = C.create_string_buffer(128)
    off, rank, shape = C.c_int64(), C.c_int(), (C.c_int64 * 4)()
    idx = 0
    while info_fn(C.byref(cfg), idx, name, C.byref(off), C.byref(rank), shape) == 0:
        out.append((name.value.decode(), int(off.value), tuple(int(shape[i]) for i in range(rank.value))))
        idx += 1
    return out


class FlatParamModule(nn.Module):
    """nn.Module whose parameters are views of ``self._flat`` (fp32) at executor-defined offsets."""

    def _init_flat(self, infos, total: int):
        self._infos = infos
        self._n_flat = int(total)
        self._flat = torch.zeros(self._n_flat, dtype=torch.float32)
        self._params_by_name = {}
        for name, off, shape in infos:
            parts = name.split(".")
            mod = self
            for p in parts[:-1]:
                if not hasattr(mod, p):
                    mod.add_module(p, nn.Module())
                mod = getattr(mod, p)
            numel = 1
            for s in shape:
                numel *= s
            param = nn.Parameter(self._flat[off:off + numel].view(shape))
            mod.register_parameter(parts[-1], param)
            self._params_by_name[name] = (param, off, numel, shape)

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._reflatten()
        return self

    def _reflatten(self):
        first = next(iter(self._params_by_name.values()))[0]
        flat = torch.zeros(self._n_flat, dtype=torch.float32, device=first.device)
        with torch.no_grad():
            for name, (param, off, numel, shape) in self._params_by_name.items():
                view = flat[off:off + numel].view(shape)
                view.copy_(param.data)
                param.data = view
        self._flat = flat

    def flat_params(self) -> torch.Tensor:
        """The flat fp32 parameter buffer; re-flattens if somebody re-pointed a parameter.

        Whole-module moves / casts go through ``_apply`` (which re-flattens) and ``load_state_dict`` copies in
        place, so the per-step check only probes a few parameters; ``flat_params(full_check=True)`` probes all."""
        return self._flat_checked(False)

    def _flat_checked(self, full_check: bool) -> torch.Tensor:
        base = self._flat.data_ptr()
        entries = self._entries()
        probe = entries if full_check else (entries[0], entries[len(entries) // 2], entries[-1])
        for param, off, numel, shape in probe:
            if param.data_ptr() != base + 4 * off or param.dtype != torch.float32:
                self._reflatten()
                break
        return self._flat

    def _entries(self):
        e = getattr(self, "_entries_cache", None)
        if e is None or len(e) != len(self._params_by_name):
            e = self._entries_cache = list(self._params_by_name.values())
            self._ordered_cache = [v[0] for v in e]
            # split plan for grad_views: parameter pieces interleaved with the alignment gaps between them
            sizes, pick, pos = [], [], 0
            for param, off, numel, shape in e:
                if off > pos:
                    sizes.append(off - pos)
                pick.append(len(sizes))
                sizes.append(numel)
                pos = off + numel
            if pos < self._n_flat:
                sizes.append(self._n_flat - pos)
            self._split_sizes, self._split_pick = sizes, pick
        return e

    def params_version_key(self):
        """Changes whenever torch writes the parameters in place: the flat block's own version counter (writes through the
        block, e.g. a broadcast) plus the counters of the parameter views (``set_data`` gave each its own; ``load_state_dict``,
        ``param.mul_`` and torch optimizers bump those).  Writes through ``param.data`` or raw pointers are invisible to it."""
        flat = self.flat_params()
        return (flat.data_ptr(), flat._version, sum(p._version for p in self.ordered_params()))

    def ordered_params(self):
        self._entries()
        return self._ordered_cache

    # ---- buffers that stay at fixed addresses across steps (the executors' CUDA-graph cache keys on them) ----
    def _take_workspace(self, batch: int, nbytes: int, device):
        pool = self.__dict__.setdefault("_ws_pool", {})
        free = pool.setdefault((batch, nbytes, str(device)), [])
        return free.pop() if free else torch.empty(nbytes, dtype=torch.uint8, device=device)

    def _return_workspace(self, batch: int, ws):
        free = self.__dict__.setdefault("_ws_pool", {}).setdefault((batch, ws.numel(), str(ws.device)), [])
        if len(free) < 2:
            free.append(ws)

    def _grad_block(self, device):
        """Zeroed flat gradient block.  The persistent block is reused only when no parameter still holds a gradient
        (i.e. after zero_grad(set_to_none=True)): with gradient accumulation the previous views must stay intact."""
        buf = getattr(self, "_grad_buf", None)
        ps = self.ordered_params()
        if buf is not None and buf.device == device and all(p.grad is None for p in ps):
            return buf.zero_()
        g = torch.zeros(self._n_flat, dtype=torch.float32, device=device)
        if buf is None or buf.device != device:
            self._grad_buf = g
        return g

    def grad_views(self, flat_grad: torch.Tensor, needs):
        """Per-parameter views of the flat gradient block (one split call + a reshape for the >1-D tensors)."""
        entries = self._entries()
        pieces = flat_grad.split_with_sizes(self._split_sizes)
        out = []
        for (param, off, numel, shape), idx, need in zip(entries, self._split_pick, needs):
            if not need:
                out.append(None)
            else:
                v = pieces[idx]
                out.append(v if len(shape) == 1 else v.view(shape))
        return out


def is_channels_last_view(t: torch.Tensor) -> bool:
    """True when a [B,C,H,W] tensor is laid out NHWC in memory (dense)."""
    return t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous()
