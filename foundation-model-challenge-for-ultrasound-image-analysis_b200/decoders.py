"""FPN decoder factory: drop-in for ``/root/reference/code/models/decoders.py``.

``build_fpn_decoder`` (decoders.py:9-60) and ``build_decoders`` (decoders.py:63-103) keep their
signatures and config keys (``model.decoder.{pyramid_channels,segmentation_channels,dropout,
merge_policy,separate_*_fpn}``); ``FPNDecoder`` keeps smp's constructor kwargs, ``out_channels``
attribute, list-of-features call convention (multitask_model.py:211) and state-dict keys
(``p5``, ``p{4,3,2}.skip_conv``, ``seg_blocks.{i}.block.{k}.block.{0,1}``).  The arithmetic runs in
``libmtus_b200.so`` (``mtus_fpn_forward/backward``); no fallback.
"""

import ctypes as C
import math
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._native import FlatParamModule, enumerate_params, precision_to_dtype, is_channels_last_view
from .encoders import default_precision


class _FpnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, nfeat, gamma, beta, *args):
        feats, params = args[:nfeat], args[nfeat:]
        needs_grad = any(ctx.needs_input_grad[2:])
        if not feats[0].is_cuda:
            raise RuntimeError("mtus_b200: the FPN decoder runs only on CUDA (sm_100a); there is no CPU fallback")
        with _lib.device_guard(feats[0]):
            out, saved = dec._run_forward(list(feats), training_plan=needs_grad, film=(gamma, beta))
        ctx.dec, ctx.saved, ctx.nfeat = dec, saved, nfeat
        ctx.film_needs = ctx.needs_input_grad[2:4]
        ctx.needs = ctx.needs_input_grad[4:]
        return out

    @staticmethod
    def backward(ctx, dout):
        dec = ctx.dec
        with _lib.device_guard(dout):
            dfeats, flat_grad, dgamma, dbeta = dec._run_backward(ctx.saved, dout, ctx.needs[:ctx.nfeat], ctx.film_needs)
        ctx.saved = None
        return (None, None, dgamma, dbeta) + tuple(dfeats) + tuple(dec.grad_views(flat_grad, ctx.needs[ctx.nfeat:]))


class FPNDecoder(FlatParamModule):
    """smp ``FPNDecoder`` (constructor exactly as used at decoders.py:42-49)."""

    def __init__(self, encoder_channels, encoder_depth=5, pyramid_channels=256, segmentation_channels=128,
                 dropout=0.2, merge_policy="add", precision: Optional[str] = None, output_dtype: Optional[str] = None,
                 channels_last_output: bool = True):
        super().__init__()
        if merge_policy not in ("add", "cat"):
            raise ValueError("`merge_policy` must be one of: ['add', 'cat'], got {}".format(merge_policy))
        if encoder_depth < 3:
            raise ValueError("Encoder depth for FPN decoder cannot be less than 3, got {}.".format(encoder_depth))
        if encoder_depth != 4 or len(encoder_channels) < 5:
            raise NotImplementedError("mtus_b200 FPN is built for the 4-stage Swin pyramid (encoder_depth=4)")
        self.out_channels = segmentation_channels if merge_policy == "add" else segmentation_channels * 4
        self.in_channels = [int(c) for c in list(encoder_channels)[-4:]]          # c2..c5
        self.pyramid_channels, self.segmentation_channels = int(pyramid_channels), int(segmentation_channels)
        self.merge_policy, self.p_drop = merge_policy, float(dropout)
        self.precision = precision or default_precision()
        self.output_dtype = output_dtype
        # the [B, C, H, W] result is laid out channels-last in memory (what the kernels produce natively and what
        # cuDNN wants for the heads' convolutions); shape and values are those of the reference's NCHW tensor
        self.channels_last_output = bool(channels_last_output)
        self.backend = _lib.BACKEND_AUTO
        cfg = self._cfg(1, [8, 4, 2, 1], True)
        L = _lib.lib()
        total = L.mtus_fpn_param_count(C.byref(cfg))
        if total <= 0:
            raise RuntimeError("mtus_b200: invalid FPN configuration (channels must be multiples of 32 / 8)")
        self._init_flat(enumerate_params(L.mtus_fpn_param_info, cfg), total)
        self.reset_parameters()

    def _cfg(self, batch, sizes, training) -> _lib.FpnConfig:
        dt, _ = precision_to_dtype(self.precision)
        c = _lib.FpnConfig()
        c.batch = batch
        for k in range(4):
            c.in_channels[k], c.sizes[k] = self.in_channels[k], sizes[k]
        c.pyramid_channels, c.seg_channels = self.pyramid_channels, self.segmentation_channels
        c.merge_cat = int(self.merge_policy == "cat")
        c.dtype, c.backend, c.training = dt, self.backend, int(training)
        return c

    def reset_parameters(self):
        """PyTorch defaults, as smp leaves them: kaiming-uniform convs, GroupNorm (1, 0)."""
        with torch.no_grad():
            for name, (p, off, numel, shape) in self._params_by_name.items():
                if p.dim() == 4:
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                elif name.endswith("block.1.weight"):
                    nn.init.ones_(p)
                elif name.endswith("block.1.bias"):
                    nn.init.zeros_(p)
                else:  # conv bias
                    fan_in = self._params_by_name[name[:-4] + "weight"][3][1]
                    bound = 1.0 / math.sqrt(fan_in)
                    nn.init.uniform_(p, -bound, bound)

    def _run_forward(self, feats: List[torch.Tensor], training_plan: bool, film=(None, None)):
        L = _lib.lib()
        dt, tdt = precision_to_dtype(self.precision)
        x0 = feats[0]
        if not x0.is_cuda:
            raise RuntimeError("mtus_b200: the FPN decoder runs only on CUDA (sm_100a); there is no CPU fallback")
        B = x0.shape[0]
        sizes = [int(f.shape[2]) for f in feats]
        for f, c in zip(feats, self.in_channels):
            if f.shape[1] != c or f.shape[2] != f.shape[3]:
                raise ValueError("feature map does not match encoder_channels / is not square")
        f32_in = all(f.dtype == torch.float32 for f in feats) and dt != _lib.F32
        want = torch.float32 if (f32_in or dt == _lib.F32) else tdt
        feats = [f if f.dtype == want else f.to(want) for f in feats]
        nhwc = all(is_channels_last_view(f) for f in feats) and not f32_in
        if not nhwc:
            feats = [f.contiguous() for f in feats]
        flat = self.flat_params()
        cfg = self._cfg(B, sizes, training_plan)
        nbytes = L.mtus_fpn_workspace_bytes(C.byref(cfg))
        if nbytes < 0:
            raise ValueError("mtus_b200: unsupported FPN geometry (each level must be exactly 2x the next)")
        # training workspaces come from a pool and return to it when their backward has run (fixed addresses for
        # the executor's graph cache); inference workspaces stay per call
        ws = self._take_workspace(B, nbytes, x0.device) if training_plan else torch.empty(nbytes, dtype=torch.uint8, device=x0.device)
        scale = drop = None
        if self.training and self.p_drop > 0.0:     # Dropout2d: whole channels, scaled by 1/(1-p)
            keep = 1.0 - self.p_drop
            scale = drop = ((torch.rand(B, self.out_channels, device=x0.device) < keep).float() / keep).contiguous()
        gamma, beta = film
        shift = None
        if gamma is not None:                       # FiLM epilogue: gamma folds into the per-(sample, channel) scale
            g = gamma.detach().float().reshape(1, self.out_channels)
            scale = (g.expand(B, -1) if drop is None else drop * g).contiguous()
        if beta is not None:
            shift = beta.detach().float().reshape(self.out_channels).contiguous()
        out_f32 = (self.output_dtype in ("fp32", "float32") or f32_in) and dt != _lib.F32
        odt = torch.float32 if (out_f32 or dt == _lib.F32) else tdt
        if self.channels_last_output:
            out = torch.empty(B, sizes[0], sizes[0], self.out_channels, dtype=odt, device=x0.device).permute(0, 3, 1, 2)
        else:
            out = torch.empty(B, self.out_channels, sizes[0], sizes[0], dtype=odt, device=x0.device)
        _lib.check(L.mtus_fpn_forward_film(C.byref(cfg), _lib.ptr_array(feats), int(nhwc), int(f32_in), _lib.ptr(flat),
                                           _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(ws), _lib.ptr(out),
                                           int(out_f32) | (2 if self.channels_last_output else 0), _lib.stream_ptr()),
                   "fpn_forward")
        saved = (cfg, ws, feats, nhwc, f32_in, flat, scale, out_f32, drop, gamma is not None) if training_plan else None
        return out, saved

    def _run_backward(self, saved, dout, feat_needs, film_needs=(False, False)):
        cfg, ws, feats, nhwc, f32_in, flat, scale, out_f32, drop, has_film = saved
        L = _lib.lib()
        dt, tdt = precision_to_dtype(self.precision)
        want_out = torch.float32 if (out_f32 or dt == _lib.F32) else tdt
        dout = dout.to(want_out)
        dout_nhwc = is_channels_last_view(dout)
        if not dout_nhwc:
            dout = dout.contiguous()
        flat_grad = self._grad_block(flat.device)
        # feature gradients live in persistent buffers: the encoder's backward converts them into its own workspace
        # right away (stream-ordered before the next step can overwrite them)
        bufs = self.__dict__.setdefault("_dfeat_bufs", {})
        dfeats = []
        for i, f in enumerate(feats):
            key = (i, tuple(f.shape), f.dtype, str(f.device), bool(nhwc))
            g = bufs.get(key)
            if g is None:
                if nhwc:   # gradients come back channels-last, like the features
                    B, Cc, H, W = f.shape
                    g = torch.empty(B, H, W, Cc, dtype=f.dtype, device=f.device).permute(0, 3, 1, 2)
                else:
                    g = torch.empty_like(f)
                bufs[key] = g
            dfeats.append(g)
        _lib.check(L.mtus_fpn_backward(C.byref(cfg), _lib.ptr_array(feats), int(nhwc), int(f32_in), _lib.ptr(flat),
                                       _lib.ptr(scale), _lib.ptr(ws), _lib.ptr(dout), int(want_out == torch.float32 and dt != _lib.F32) | (2 if dout_nhwc else 0),
                                       _lib.ptr_array(dfeats), int(nhwc), int(f32_in), _lib.ptr(flat_grad), _lib.stream_ptr()),
                   "fpn_backward")
        dgamma = dbeta = None
        if has_film and any(film_needs):
            # dgamma[c] = sum dout * drop[b,c] * merged, dbeta[c] = sum dout: one pass over dout and the four tower outputs
            # (still in the workspace); needs the gradient channels-last, which is what the heads hand back
            if not dout_nhwc:
                dout = dout.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
            es = 2 if dt == _lib.BF16 else 4
            S2, Sc = int(feats[0].shape[2]), self.segmentation_channels
            srcs = []
            for lvl in range(4):
                off = L.mtus_fpn_tower_output_offset(C.byref(cfg), lvl)
                srcs.append(ws[off:off + cfg.batch * S2 * S2 * Sc * es])
            dgamma = torch.zeros(self.out_channels, dtype=torch.float32, device=flat.device)
            dbeta = torch.zeros_like(dgamma)
            _lib.check(L.mtus_film_grad(_lib.ptr(dout), int(want_out == torch.float32 and dt != _lib.F32), _lib.ptr_array(srcs), 4,
                                        int(self.merge_policy == "cat"), _lib.ptr(drop), _lib.ptr(dgamma), _lib.ptr(dbeta), cfg.batch,
                                        S2 * S2, Sc, dt, _lib.stream_ptr()), "film_grad")
            dgamma = dgamma if film_needs[0] else None
            dbeta = dbeta if film_needs[1] else None
        self._last_flat_grad = flat_grad
        self._return_workspace(cfg.batch, ws)
        return [g if n else None for g, n in zip(dfeats, feat_needs)], flat_grad, dgamma, dbeta

    def forward(self, features: List[torch.Tensor], film=None) -> torch.Tensor:
        """``film = (gamma [C_out], beta [C_out] | None)`` applies the reference's FiLM modulation of the decoder output
        (film_layer.py:94-99) inside the merge kernel -- no extra pass over the [B, C_out, H/4, W/4] map."""
        feats = list(features)[-4:]
        gamma, beta = film if film is not None else (None, None)
        return _FpnFn.apply(self, len(feats), gamma, beta, *feats, *self.ordered_params())


def build_fpn_decoder(encoder, config, decoder_type="seg", precision: Optional[str] = None,
                      output_dtype: Optional[str] = None):
    """Drop-in for ``build_fpn_decoder`` (decoders.py:9-60)."""
    pyramid_channels = int(config.get("model.decoder.pyramid_channels", 256))
    segmentation_channels = int(config.get("model.decoder.segmentation_channels", 128))
    dropout = float(config.get("model.decoder.dropout", 0.2))
    merge_policy = config.get("model.decoder.merge_policy", "cat")
    if hasattr(encoder, "is_timm_encoder") and encoder.is_timm_encoder:
        encoder_channels = encoder.out_channels
    else:
        encoder_channels = [3] + list(encoder.out_channels)
    encoder_depth = len(encoder_channels) - 1
    if precision is None:
        precision = getattr(getattr(encoder, "model", None), "precision", None)
        if precision is None:
            mp = config.get("device.mixed_precision", None)
            precision = default_precision() if mp is None else ("bf16" if mp else "fp32")
    decoder = FPNDecoder(encoder_channels=encoder_channels, encoder_depth=encoder_depth,
                         pyramid_channels=pyramid_channels, segmentation_channels=segmentation_channels,
                         dropout=dropout, merge_policy=merge_policy, precision=precision, output_dtype=output_dtype)
    suffix = {"seg": "segmentation", "det": "detection", "cls": "classification", "reg": "regression"}.get(decoder_type, decoder_type)
    print(f"Built FPN decoder for {suffix}")
    return decoder


def build_decoders(encoder, config, precision: Optional[str] = None, output_dtype: Optional[str] = None):
    """Drop-in for ``build_decoders`` (decoders.py:63-103): same keys, same aliasing when not separate."""
    decoders = {"fpn_seg": build_fpn_decoder(encoder, config, "seg", precision, output_dtype)}
    for key, flag, what in (("fpn_det", "separate_detection_fpn", "detection"),
                            ("fpn_cls", "separate_classification_fpn", "classification"),
                            ("fpn_reg", "separate_regression_fpn", "regression")):
        if config.get("model.decoder." + flag, True):
            decoders[key] = build_fpn_decoder(encoder, config, key[4:], precision, output_dtype)
            print(f"Using separate FPN decoder for {what}")
        else:
            decoders[key] = decoders["fpn_seg"]
            print(f"Sharing FPN decoder between segmentation and {what}")
    return decoders
