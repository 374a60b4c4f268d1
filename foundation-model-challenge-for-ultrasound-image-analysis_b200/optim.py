"""AdamW + global-norm gradient clipping over the flat parameter blocks of the native modules.

Same update rule and step semantics as the reference's ``torch.optim.AdamW`` + ``clip_grad_norm_``
(``/root/reference/code/train.py:176-219, 446, 455``): grouped learning rates (encoder x0.1, heads x1.0), decoupled
weight decay, parameters that received no gradient this step are skipped entirely (no decay, no moment update).
The Swin encoder and each FPN decoder own ONE contiguous fp32 parameter block and ONE gradient block, so their update is
a single kernel each (``mtus_adamw_flat``) instead of a multi-tensor sweep over ~330 tensors; the clip coefficient
stays on the device.  Everything that is not a flat module (the PyTorch task heads) goes through torch's own AdamW.

``FlatAdamW`` IS a ``torch.optim.Optimizer``: ``param_groups`` holds one group per flat block followed by the groups of
the remaining parameters, each with ``lr`` / ``initial_lr``-capable entries, and ``step()`` reads the learning rate from
the group -- so the reference's schedulers (``CosineAnnealingLR`` / ``StepLR`` / ``ReduceLROnPlateau``,
``code/train.py:222-253``, stepped at ``:698-704``) attach to it unchanged.
"""

import contextlib
import os
from typing import Dict, List

import torch

from . import _lib
from ._native import FlatParamModule


# MTUS_OPT_SHADOW=0: the optimizer leaves the encoder's bf16 parameter shadow alone and every forward re-casts it
_OPT_WRITES_SHADOW = os.environ.get("MTUS_OPT_SHADOW", "1") != "0"


class FlatAdamW(torch.optim.Optimizer):
    handles_clipping = True

    def __init__(self, model, lr: float = 1e-4, weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 encoder_lr_multiplier: float = 0.1, head_lr_multiplier: float = 1.0, max_grad_norm: float = 1.0):
        self.model = model
        self.max_norm = float(max_grad_norm)
        enc_params, head_params = model.get_trainable_parameters()
        enc_ids = {id(p) for p in enc_params}
        self.flat: List[Dict] = []
        self._all_params = None
        flat_ids = set()
        groups = []
        for mod in model.modules():
            if isinstance(mod, FlatParamModule):
                ps = mod.ordered_params()
                if not all(p.requires_grad for p in ps):
                    continue                                 # partially frozen block: leave it to torch
                is_enc = all(id(p) in enc_ids for p in ps)
                self.flat.append({"module": mod, "m": None, "v": None, "step": 0})
                groups.append({"params": list(ps), "lr": lr * (encoder_lr_multiplier if is_enc else head_lr_multiplier),
                               "flat_block": len(self.flat) - 1})
                flat_ids |= {id(p) for p in ps}
        rest_enc = [p for p in enc_params if id(p) not in flat_ids]
        rest_head = [p for p in head_params if id(p) not in flat_ids]
        rest_groups = []
        if rest_enc:
            rest_groups.append({"params": rest_enc, "lr": lr * encoder_lr_multiplier})
        if rest_head:
            rest_groups.append({"params": rest_head, "lr": lr * head_lr_multiplier})
        self.rest_params = rest_enc + rest_head
        self._n_flat_groups = len(groups)
        defaults = dict(lr=lr, betas=tuple(betas), eps=float(eps), weight_decay=float(weight_decay), flat_block=-1)
        super().__init__(groups + [dict(g) for g in rest_groups], defaults)
        # the remaining parameters are stepped by torch's own (fused) AdamW; its groups mirror ours and take their
        # hyper-parameters from ours right before every step, so a scheduler acting on self.param_groups drives both
        self.torch_opt = torch.optim.AdamW(rest_groups, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps,
                                           fused=all(p.is_cuda for p in self.rest_params)) if rest_groups else None

    # -- torch.optim.Optimizer surface -----------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        if self._all_params is None:
            self._all_params = list(self.model.parameters())
        for p in self._all_params:
            p.grad = None
        for f in self.flat:
            f["module"]._last_flat_grad = None

    def _flat_grad(self, f):
        g = getattr(f["module"], "_last_flat_grad", None)
        if g is None:
            return None
        # the module handed views of this buffer to autograd as .grad; somebody may have replaced them
        first = f["module"].ordered_params()[0]
        if first.grad is None or first.grad.data_ptr() < g.data_ptr() or first.grad.data_ptr() >= g.data_ptr() + g.numel() * 4:
            return None
        return g

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        dev = next(self.model.parameters()).device
        with (torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()):
            st = _lib.stream_ptr(dev) if dev.type == "cuda" else None
            active = [(f, self._flat_grad(f)) for f in self.flat]
            fallback = [f for f, g in active if g is None and any(p.grad is not None for p in f["module"].ordered_params())]
            if fallback:
                raise RuntimeError("FlatAdamW: a flat module's gradients are not views of its gradient block")
            active = [(i, f, g) for i, (f, g) in enumerate(active) if g is not None]
            rest_grads = [p.grad for p in self.rest_params if p.grad is not None]
            scale = None
            if self.max_norm > 0:
                sq = torch.zeros(1, dtype=torch.float32, device=dev)
                for _, f, g in active:
                    _lib.check(L.mtus_sumsq(_lib.ptr(g), g.numel(), _lib.ptr(sq), st), "sumsq")
                if rest_grads:
                    norms = torch._foreach_norm(rest_grads)
                    sq += torch.stack([n.float() for n in norms]).square().sum()
                total = sq.sqrt()
                scale = (self.max_norm / (total + 1e-6)).clamp(max=1.0)        # clip_grad_norm_ coefficient, on device
                if rest_grads:
                    torch._foreach_mul_(rest_grads, scale.squeeze(0))
            for i, f, g in active:
                grp = self.param_groups[i]
                mod = f["module"]
                p = mod.flat_params()
                if f["m"] is None:
                    f["m"], f["v"] = torch.zeros_like(p), torch.zeros_like(p)
                if f["m"].device != p.device or f["m"].numel() != p.numel():
                    raise RuntimeError("FlatAdamW: optimizer state and parameter block disagree in device or size "
                                       "(load_state_dict moves the moments to the parameters' device; did the model move afterwards?)")
                f["step"] += 1
                b1, b2 = grp["betas"]
                lp = getattr(mod, "_lp_buf", None) if _OPT_WRITES_SHADOW else None
                if lp is not None and lp.device == p.device and lp.numel() == p.numel() and lp.dtype == torch.bfloat16:
                    # the encoder's bf16 operand shadow is refreshed by the update itself; the module skips its own cast
                    # pass while the parameter block stays untouched (same tensor, same version counter)
                    _lib.check(L.mtus_adamw_flat_shadow(_lib.ptr(p), _lib.ptr(g), _lib.ptr(f["m"]), _lib.ptr(f["v"]), _lib.ptr(lp), p.numel(),
                                                        float(grp["lr"]), float(b1), float(b2), float(grp["eps"]), float(grp["weight_decay"]),
                                                        f["step"], _lib.ptr(scale), st), "adamw_flat_shadow")
                    mod._lp_fresh_key = mod.params_version_key() + (lp.data_ptr(),)
                else:
                    _lib.check(L.mtus_adamw_flat(_lib.ptr(p), _lib.ptr(g), _lib.ptr(f["m"]), _lib.ptr(f["v"]), p.numel(), float(grp["lr"]),
                                                 float(b1), float(b2), float(grp["eps"]), float(grp["weight_decay"]), f["step"],
                                                 _lib.ptr(scale), st), "adamw_flat")
            if self.torch_opt is not None:
                for mine, theirs in zip(self.param_groups[self._n_flat_groups:], self.torch_opt.param_groups):
                    for k in ("lr", "betas", "eps", "weight_decay"):
                        theirs[k] = mine[k]
                self.torch_opt.step()
        return loss

    def state_dict(self):
        return {"flat": [{"m": f["m"], "v": f["v"], "step": f["step"]} for f in self.flat],
                "groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "torch": self.torch_opt.state_dict() if self.torch_opt is not None else None}

    def load_state_dict(self, sd):
        if len(sd["flat"]) != len(self.flat):
            raise ValueError("FlatAdamW.load_state_dict: number of flat blocks differs")
        for f, s in zip(self.flat, sd["flat"]):
            p = f["module"].flat_params()
            for key in ("m", "v"):
                t = s[key]
                if t is not None:
                    if t.numel() != p.numel():
                        raise ValueError("FlatAdamW.load_state_dict: moment size does not match the parameter block")
                    t = t.to(device=p.device, dtype=p.dtype).clone()
                f[key] = t
            f["step"] = int(s["step"])
        for g, s in zip(self.param_groups, sd.get("groups", [])):
            g.update({k: v for k, v in s.items() if k != "params"})
        if self.torch_opt is not None and sd.get("torch") is not None:
            self.torch_opt.load_state_dict(sd["torch"])
