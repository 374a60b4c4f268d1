"""Swin encoder factory: drop-in for ``/root/reference/code/models/encoders.py`` (swin branch).

``build_encoder(config, task_ids)`` (encoders.py:665-691), ``SwinTransformerEncoder``
(encoders.py:37-160) and ``SWIN_MODEL_MAPPING`` (encoders.py:14-19) keep their names, arguments,
attributes (``is_timm_encoder``, ``out_channels``, ``output_stride``, ``supports_task_id``,
``handles_moe``, ``use_moe``, ``get_moe_aux_loss``, ``get_moe_stats``) and the ``encoder.model.*``
state-dict keys of timm's FeatureListNet, but the arithmetic runs in the sm_100a kernels of
``libmtus_b200.so`` through ONE C-ABI call per direction (``mtus_swin_forward/backward``).
There is no PyTorch / CPU fallback: on a box without the library or without CUDA the forward raises.
"""

import ctypes as C
import math
import os
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._native import FlatParamModule, enumerate_params, precision_to_dtype, is_channels_last_view

# code/models/encoders.py:14-19
SWIN_MODEL_MAPPING = {
    "swin_t": "swin_tiny_patch4_window7_224",
    "swin_s": "swin_small_patch4_window7_224",
    "swin_b": "swin_base_patch4_window7_224",
    "swin_l": "swin_large_patch4_window7_224",
}

# timm model registry entries reachable through the reference (name -> embed_dim, depths, heads, window)
SWIN_ARCHS = {
    "swin_tiny_patch4_window7_224": (96, (2, 2, 6, 2), (3, 6, 12, 24), 7),
    "swin_small_patch4_window7_224": (96, (2, 2, 18, 2), (3, 6, 12, 24), 7),
    "swin_base_patch4_window7_224": (128, (2, 2, 18, 2), (4, 8, 16, 32), 7),
    "swin_large_patch4_window7_224": (192, (2, 2, 18, 2), (6, 12, 24, 48), 7),
    "swin_base_patch4_window12_384": (128, (2, 2, 18, 2), (4, 8, 16, 32), 12),
    "swin_large_patch4_window12_384": (192, (2, 2, 18, 2), (6, 12, 24, 48), 12),
    "swin_micro_patch4_window7_test": (32, (2, 2, 2, 2), (1, 2, 4, 8), 7),  # test-only (see oracle/swin.py)
}

# hook used by parallel.GradAllReducer: called as fn(flat_grad, lo, hi) after each backward stage chunk
_STAGE_GRAD_HOOK = None
# hook used by parallel.GradAllReducer: called as fn() when the encoder's backward starts (heads / decoders are done)
_PRE_BACKWARD_HOOK = None


def default_precision() -> str:
    return os.environ.get("MTUS_PRECISION", "bf16")


class _FeatureInfo:
    def __init__(self, chans, reds):
        self._c, self._r = list(chans), list(reds)

    def channels(self):
        return list(self._c)

    def reduction(self):
        return list(self._r)


class _SwinFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, core, x, *params):
        # a trainable module in front of the encoder (TaskPrompt2D) needs the backward plan even with a frozen encoder
        needs_grad = any(ctx.needs_input_grad[1:])
        if not x.is_cuda:
            raise RuntimeError("mtus_b200: the Swin encoder runs only on CUDA (sm_100a); there is no CPU fallback")
        with _lib.device_guard(x):
            feats, saved = core._run_forward(x, training_plan=needs_grad)
        ctx.set_materialize_grads(False)         # unused features arrive as None, not as zero tensors to convert
        ctx.core = core
        ctx.saved = saved
        ctx.param_needs = ctx.needs_input_grad[2:]
        return tuple(feats)

    @staticmethod
    def backward(ctx, *dfeats):
        core = ctx.core
        with _lib.device_guard(core._flat):          # autograd's backward thread may sit on another device
            flat_grad, dx = core._run_backward(ctx.saved, dfeats, want_input_grad=ctx.needs_input_grad[1])
        ctx.saved = None
        return (None, dx) + tuple(core.grad_views(flat_grad, ctx.param_needs))


class SwinCore(FlatParamModule):
    """Stands where timm's ``FeatureListNet(SwinTransformer)`` stands (``encoder.model``)."""

    def __init__(self, name: str, img_size: int = 224, drop_path_rate: float = 0.1, precision: Optional[str] = None,
                 output_dtype: Optional[str] = None, zero_copy_features: bool = False):
        super().__init__()
        if name not in SWIN_ARCHS:
            raise RuntimeError(f"Unknown model ({name})")
        self.arch_name = name
        self.embed_dim, self.depths, self.num_heads, self.window = SWIN_ARCHS[name]
        self.img_size = int(img_size)
        if self.img_size % 4:
            raise ValueError("image size must be a multiple of the patch size 4")
        self.precision = precision or default_precision()
        self.output_dtype = output_dtype          # None -> activation dtype; 'fp32' -> fp32 NCHW features
        self.zero_copy_features = zero_copy_features
        self.drop_path_rate = float(drop_path_rate)
        # Normalize(mean, std, max_pixel_value=255) of the reference's input pipeline (code/train.py:35-44): applied by the
        # kernels when forward() is handed the raw uint8 [B,H,W,3] batch; ImageNet statistics by default
        self.input_mean, self.input_std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
        self.backend = _lib.BACKEND_AUTO
        nblk = sum(self.depths)
        self._dpr = torch.linspace(0, self.drop_path_rate, nblk).tolist()   # timm: linspace(0, rate, sum(depths))
        chans = [self.embed_dim * 2 ** i for i in range(4)]
        self.feature_info = _FeatureInfo(chans, [4, 8, 16, 32])
        res, r = [], self.img_size // 4
        for i in range(4):
            if i > 0:
                r = (r + 1) // 2
            res.append(r)
        self.resolutions = res
        cfg = self._cfg(1, True)
        L = _lib.lib()
        total = L.mtus_swin_param_count(C.byref(cfg))
        if total <= 0:
            raise RuntimeError("mtus_b200: invalid Swin configuration")
        self._init_flat(enumerate_params(L.mtus_swin_param_info, cfg), total)
        self._stage_slices = self._compute_stage_slices()
        self.reset_parameters()

    # -- configuration ---------------------------------------------------------------------------
    def _cfg(self, batch: int, training: bool) -> _lib.SwinConfig:
        dt, _ = precision_to_dtype(self.precision)
        c = _lib.SwinConfig()
        c.batch, c.img_size, c.embed_dim, c.window = batch, self.img_size, self.embed_dim, self.window
        for i in range(4):
            c.depths[i], c.heads[i] = self.depths[i], self.num_heads[i]
        c.dtype, c.backend, c.training, c.ln_eps = dt, self.backend, int(training), 1e-5
        return c

    def _compute_stage_slices(self):
        """[lo, hi) element ranges of the flat buffer: patch_embed, layers_0 .. layers_3."""
        bounds = {}
        for name, (p, off, numel, shape) in self._params_by_name.items():
            key = name.split(".")[0]
            lo, hi = bounds.get(key, (off, off))
            bounds[key] = (min(lo, off), max(hi, off + ((numel + 7) // 8) * 8))
        return [bounds["patch_embed"]] + [bounds[f"layers_{i}"] for i in range(4)]

    def reset_parameters(self):
        """timm init: trunc_normal(.02) Linear weights + rel-pos tables, zero biases, LN (1, 0); conv default."""
        with torch.no_grad():
            for name, (p, off, numel, shape) in self._params_by_name.items():
                if name == "patch_embed.proj.weight":
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                elif name == "patch_embed.proj.bias":
                    bound = 1.0 / math.sqrt(48)
                    nn.init.uniform_(p, -bound, bound)
                elif "norm" in name.split(".")[-2]:
                    (nn.init.ones_ if name.endswith("weight") else nn.init.zeros_)(p)
                elif name.endswith("relative_position_bias_table") or name.endswith(".weight"):
                    nn.init.trunc_normal_(p, std=0.02)
                else:
                    nn.init.zeros_(p)

    # -- execution -------------------------------------------------------------------------------
    def _droppath_scales(self, batch: int, device):
        if not self.training or self.drop_path_rate <= 0.0:
            return None
        keep = 1.0 - torch.tensor(self._dpr, device=device, dtype=torch.float32).repeat_interleave(2)  # [2*nblk]
        mask = torch.rand(keep.numel(), batch, device=device) < keep[:, None]
        return (mask.float() / keep[:, None]).contiguous()

    def _run_forward(self, x: torch.Tensor, training_plan: bool):
        if not x.is_cuda:
            raise RuntimeError("mtus_b200: the Swin encoder runs only on CUDA (sm_100a); there is no CPU fallback")
        x_u8 = x.dtype == torch.uint8
        if x_u8:                                   # raw HWC batch: normalisation fused into the patch-embed operand
            B, H, W, Cin = x.shape
        else:
            B, Cin, H, W = x.shape
        if Cin != 3:
            raise ValueError("expected a 3-channel image")
        # timm strict image size (encoders.py:58 threads img_size through for this reason)
        assert H == self.img_size, f"Input height ({H}) doesn't match model ({self.img_size})."
        assert W == self.img_size, f"Input width ({W}) doesn't match model ({self.img_size})."
        L = _lib.lib()
        dt, tdt = precision_to_dtype(self.precision)
        flat = self.flat_params()
        if flat.device != x.device:
            raise RuntimeError("mtus_b200: encoder parameters and input live on different devices")
        x_is_f32 = x.dtype == torch.float32
        if not x_u8 and not x_is_f32 and x.dtype != tdt:
            x = x.to(tdt)
        x = x.contiguous()
        cfg = self._cfg(B, training_plan)
        # Buffers the executor's CUDA-graph cache keys on live at fixed addresses: the bf16 parameter shadow is one
        # persistent buffer; training workspaces (saved for backward) come from a small pool and go back to it when
        # their backward has run.  Inference workspaces stay per call (the zero-copy features are views of them).
        ws_bytes = L.mtus_swin_workspace_bytes(C.byref(cfg))
        ws = self._take_workspace(B, ws_bytes, x.device) if training_plan else torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        lp = None
        if dt == _lib.BF16:
            lp = getattr(self, "_lp_buf", None)
            if lp is None or lp.device != x.device or lp.numel() != self._n_flat:
                lp = self._lp_buf = torch.empty(self._n_flat, dtype=torch.bfloat16, device=x.device)
            # FlatAdamW.step writes the shadow together with the update (mtus_adamw_flat_shadow) and records the parameter
            # block it is valid for; any later in-place change of the parameters through torch (load_state_dict, broadcast,
            # another optimizer) bumps the block's version counter and the cast below runs again.  Only trusted in training
            # mode: writes through ``param.data`` bypass the counter, and those belong to evaluation-time weight swaps.
            if not (self.training and getattr(self, "_lp_fresh_key", None) == self.params_version_key() + (lp.data_ptr(),)):
                _lib.check(L.mtus_cast_f32_to_bf16(_lib.ptr(flat), _lib.ptr(lp), self._n_flat, _lib.stream_ptr()), "cast")
                self._lp_fresh_key = None
        dp = self._droppath_scales(B, x.device)
        out_f32 = self.output_dtype in ("fp32", "float32") and dt != _lib.F32
        feats, feat_ptrs = [], [None] * 4
        if self.zero_copy_features and not out_f32:
            pass
        else:
            for i in range(4):
                r, ch = self.resolutions[i], self.embed_dim * 2 ** i
                feats.append(torch.empty(B, ch, r, r, dtype=torch.float32 if out_f32 else tdt, device=x.device))
            feat_ptrs = feats
        if x_u8:
            mean3, std3 = (C.c_float * 3)(*self.input_mean), (C.c_float * 3)(*self.input_std)
            _lib.check(L.mtus_swin_forward_u8(C.byref(cfg), _lib.ptr(x), mean3, std3, _lib.ptr(flat), _lib.ptr(lp), _lib.ptr(dp),
                                              _lib.ptr(ws), _lib.ptr_array(feat_ptrs), 0, int(out_f32), _lib.stream_ptr()),
                       "swin_forward_u8")
        else:
            _lib.check(L.mtus_swin_forward(C.byref(cfg), _lib.ptr(x), int(x_is_f32), _lib.ptr(flat), _lib.ptr(lp), _lib.ptr(dp),
                                           _lib.ptr(ws), _lib.ptr_array(feat_ptrs), 0, int(out_f32), _lib.stream_ptr()),
                       "swin_forward")
        if not feats:   # zero-copy: channels-last views of the stage outputs inside the workspace
            es = 2 if dt == _lib.BF16 else 4
            for i in range(4):
                r, ch = self.resolutions[i], self.embed_dim * 2 ** i
                off = L.mtus_swin_feature_offset(C.byref(cfg), i)
                v = ws[off:off + B * r * r * ch * es].view(tdt).view(B, r, r, ch).permute(0, 3, 1, 2)
                feats.append(v)
        saved = (cfg, ws, lp, dp, flat, out_f32) if training_plan else None
        return feats, saved

    def _run_backward(self, saved, dfeats, want_input_grad=False):
        cfg, ws, lp, dp, flat, out_f32 = saved
        L = _lib.lib()
        dt, tdt = precision_to_dtype(self.precision)
        flat_grad = self._grad_block(flat.device)
        want = torch.float32 if (out_f32 or dt == _lib.F32) else tdt
        gs = [None if g is None else g.to(want) for g in dfeats]
        layouts = [is_channels_last_view(g) for g in gs if g is not None]
        nhwc = bool(layouts) and all(layouts) and not out_f32
        gs = [None if g is None else (g if (nhwc or g.is_contiguous()) else g.contiguous()) for g in gs]
        hook = _STAGE_GRAD_HOOK
        if _PRE_BACKWARD_HOOK is not None:
            _PRE_BACKWARD_HOOK()
        nblk = sum(self.depths)
        chunks = self._backward_chunks() if hook is not None else [(nblk, 0, 0, self._n_flat)]
        for b_hi, b_lo, s_lo, s_hi in chunks:
            _lib.check(L.mtus_swin_backward_blocks(C.byref(cfg), _lib.ptr(flat), _lib.ptr(lp), _lib.ptr(dp), _lib.ptr(ws),
                                                   _lib.ptr_array(gs), int(nhwc), int(want == torch.float32 and dt != _lib.F32),
                                                   _lib.ptr(flat_grad), b_hi, b_lo, _lib.stream_ptr()), "swin_backward")
            if hook is not None:
                hook(flat_grad, s_lo, s_hi)
        self._last_flat_grad = flat_grad
        dx = None
        if want_input_grad:       # d(loss)/d(image) for the module in front of the encoder (mtus_swin_input_grad)
            dx = torch.empty(cfg.batch, 3, self.img_size, self.img_size, dtype=torch.float32, device=flat.device)
            _lib.check(L.mtus_swin_input_grad(C.byref(cfg), _lib.ptr(flat), _lib.ptr(lp), _lib.ptr(ws), _lib.ptr(dx),
                                              _lib.stream_ptr()), "swin_input_grad")
        self._return_workspace(cfg.batch, ws)
        return flat_grad, dx

    def _backward_chunks(self, blocks_per_chunk: int = 0):
        """[(block_hi, block_lo, grad_lo, grad_hi)] in backward order: runs of ``blocks_per_chunk`` blocks that never
        cross a stage, each with the slice of the flat gradient it completes (the chunk holding a stage's block 0 also
        completes that stage's PatchMerging -- or, for stage 0, the patch-embed -- gradients, which precede the
        blocks in the parameter layout).  Used by the data-parallel wrapper to start all-reducing early."""
        # default: one chunk per stage.  Measured at 2 GPUs (profiles/r2_dp_chunk_sweep.txt): 6 blocks per chunk 11.75 ms/step,
        # 12 -> 11.48, per stage -> 11.27 (one GPU: 10.88): every extra chunk costs a graph launch, a join of the weight-gradient
        # stream and an NCCL enqueue, which outweighs starting the reduction of the long stage earlier
        k = blocks_per_chunk or int(os.environ.get("MTUS_DP_BLOCKS_PER_CHUNK", "24"))
        cached = getattr(self, "_chunk_cache", None)
        if cached is not None and cached[0] == k:
            return cached[1]
        first_off = {}
        for name, (p, off, numel, shape) in self._params_by_name.items():
            parts = name.split(".")
            if parts[0].startswith("layers_") and parts[1] == "blocks":
                key = (int(parts[0][7:]), int(parts[2]))
                first_off[key] = min(first_off.get(key, off), off)
        chunks, g_end = [], sum(self.depths)
        for i in (3, 2, 1, 0):
            d, first = self.depths[i], g_end - self.depths[i]
            stage_lo, stage_hi = self._stage_slices[i + 1]
            if i == 0:
                stage_lo = self._stage_slices[0][0]           # patch_embed gradients finish with stage 0
            j_hi = d
            while j_hi > 0:
                j_lo = max(0, j_hi - k)
                lo = stage_lo if j_lo == 0 else first_off[(i, j_lo)]
                hi = stage_hi if j_hi == d else first_off[(i, j_hi)]
                chunks.append((first + j_hi, first + j_lo, lo, hi))
                j_hi = j_lo
            g_end = first
        self._chunk_cache = (k, chunks)
        return chunks

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        """Returns the four stage outputs as [B,C,H,W] tensors (strides 4/8/16/32).  ``x``: the normalised [B,3,H,W] batch
        (fp32 / bf16, what the reference's DataLoader yields) or the raw uint8 [B,H,W,3] batch, normalised on the fly."""
        return list(_SwinFn.apply(self, x, *self.ordered_params()))


class SwinTransformerEncoder(nn.Module):
    """Same constructor / attributes as the reference wrapper (encoders.py:37-160)."""
    is_timm_encoder = True

    def __init__(self, model_name: str = "swin_b", pretrained: bool = True, img_size: int = 224,
                 moe_config: Optional[dict] = None, task_ids: Optional[List[str]] = None,
                 precision: Optional[str] = None, output_dtype: Optional[str] = None,
                 zero_copy_features: bool = False, drop_path_rate: float = 0.1, pretrained_path: Optional[str] = None):
        super().__init__()
        full_model_name = SWIN_MODEL_MAPPING.get(model_name, model_name)
        self.model = SwinCore(full_model_name, img_size=img_size, drop_path_rate=drop_path_rate, precision=precision,
                              output_dtype=output_dtype, zero_copy_features=zero_copy_features)
        if pretrained:
            # timm.create_model(..., pretrained=True) (encoders.py:53-59) downloads; here the weights come from a local
            # timm / torchvision / reference checkpoint file (checkpoint.py), never from the network
            from . import checkpoint as _ckpt
            path = _ckpt.resolve_pretrained_path(full_model_name, pretrained_path)
            if path is None:
                raise RuntimeError(
                    f"mtus_b200: pretrained weights for {full_model_name!r} requested but no local file was found (no "
                    f"network here). Put {full_model_name}.pth (timm or torchvision state dict) into $MTUS_PRETRAINED_DIR, "
                    "pass model.encoder.pretrained: /path/to/file, or set model.encoder.pretrained: null.")
            _ckpt.load_pretrained(self.model, path)
            self.pretrained_path = path
        self._out_channels = self.model.feature_info.channels()
        self.output_stride = 32
        moe_cfg = moe_config or {}
        self.use_moe = bool(moe_cfg.get("enabled", False))
        if self.use_moe:
            raise NotImplementedError("mtus_b200: MoE blocks are outside the hot path (SURVEY §2 row 9); disable model.moe")
        self.moe_stage_indices = moe_cfg.get("stage_indices", None)
        self.supports_task_id = False
        self.handles_moe = False

    def forward(self, x, task_id=None):
        # the kernels already emit NCHW (or channels-last views): no permute().contiguous() pass (encoders.py:106)
        return self.model(x)

    def get_moe_aux_loss(self):
        return torch.tensor(0.0, device=self.model.flat_params().device)

    def get_moe_stats(self):
        return []

    @property
    def out_channels(self):
        return [3] + list(self._out_channels)


def build_encoder(config, task_ids=None, precision: Optional[str] = None, output_dtype: Optional[str] = None,
                  zero_copy_features: bool = False):
    """Drop-in for ``build_encoder`` (encoders.py:665-691); only the ``swin_`` branch is in scope."""
    encoder_name = config.get("model.encoder.name")
    encoder_weights = config.get("model.encoder.pretrained")
    img_size = config.get("data.image_size", 224)
    if not encoder_name.startswith("swin_"):
        raise NotImplementedError(f"mtus_b200 builds only the Swin encoders (got {encoder_name!r}); SURVEY §2 row 7")
    pretrained = (encoder_weights == "imagenet" or encoder_weights is not None)
    if precision is None:
        mp = config.get("device.mixed_precision", None)
        precision = default_precision() if mp is None else ("bf16" if mp else "fp32")
    encoder = SwinTransformerEncoder(model_name=encoder_name, pretrained=pretrained, img_size=img_size,
                                     moe_config=config.get("model.moe", {}), task_ids=task_ids, precision=precision,
                                     output_dtype=output_dtype, zero_copy_features=zero_copy_features,
                                     pretrained_path=encoder_weights if isinstance(encoder_weights, str) else None)
    mean, std = config.get("data.augmentation.normalize.mean", None), config.get("data.augmentation.normalize.std", None)
    if mean is not None and std is not None:
        encoder.model.input_mean, encoder.model.input_std = tuple(float(v) for v in mean), tuple(float(v) for v in std)
    print(f"Loaded Swin Transformer: {encoder_name} (img_size={img_size}, precision={precision}, sm_100a kernels)")
    return encoder
