"""Task heads used to drive the hot path.

The conv stacks that hold the heads' FLOPs run on the library's engines: the segmentation head's
Conv3x3(512->128) -> GroupNorm -> SiLU x2 and the baseline detection head's Conv3x3 -> BatchNorm2d -> ReLU x2 use the
implicit-GEMM tcgen05 convolution (``mtus_conv3x3_{fwd,dgrad,wgrad}``) and the fused normalise + activation kernels
(``mtus_groupnorm_act_*`` / ``mtus_batchnorm_act_*``) on channels-last activations; what remains in PyTorch is the final
1x1 / k x k projection to 2-5 channels, the bilinear upsampling of the logits and the pooled Linear heads.

Only the head types selected by ``configs/swin_b.yaml`` are provided (the "baseline" family of
``/root/reference/code/models/heads.py``): the standard segmentation head (heads.py:16-42, reached
because ``type: baseline`` falls through at heads.py:478-496), ``BaselineSmpClassificationHead``
(:363-385), ``BaselineRegressionHead`` (:388-407) and ``BaselineFPNGridDetectionHead`` (:410-434).
Module / parameter names match the reference so its checkpoints load.  Under the reference's own
``MultiTaskModel`` the reference's heads are used instead; these exist because /root/reference (and
smp) are not available on the benchmark box.
"""

import os

import torch
import torch.nn as nn


def _gn_groups(channels: int) -> int:
    groups = min(32, channels)
    while channels % groups != 0:
        groups -= 1
    return groups


class _SmpSegmentationHead(nn.Sequential):
    """Conv2d(k, pad k//2) -> bilinear upsampling (align_corners=True) -> Identity (smp base head)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, upsampling=1):
        up = nn.UpsamplingBilinear2d(scale_factor=upsampling) if upsampling > 1 else nn.Identity()
        super().__init__(nn.Conv2d(in_channels, out_channels, kernel_size, padding=kernel_size // 2), up, nn.Identity())


class _SmpClassificationHead(nn.Sequential):
    """GAP -> Flatten -> Dropout -> Linear -> Identity (smp base head)."""

    def __init__(self, in_channels, classes, dropout=0.2):
        drop = nn.Dropout(p=dropout, inplace=True) if dropout else nn.Identity()
        super().__init__(nn.AdaptiveAvgPool2d(1), nn.Flatten(), drop, nn.Linear(in_channels, classes), nn.Identity())


class _GroupNormSiLUFn(torch.autograd.Function):
    """silu(group_norm(x)) for a channels-last CUDA tensor through the library's GroupNorm kernels (one statistics pass,
    one fused normalise + SiLU pass; backward recomputes the pre-activation) -- replaces nn.GroupNorm + nn.SiLU, which
    under autocast run in fp32 behind two dtype / layout copies of the [B, 128, 56, 56] activation each way."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps):
        from . import ops
        xn = x.permute(0, 2, 3, 1)                       # NHWC view of a channels-last tensor
        if not xn.is_contiguous():
            xn = xn.contiguous()
        w, b = weight.float().contiguous(), bias.float().contiguous()
        y, mean, rstd = ops.groupnorm_silu_fwd(xn, w, b, groups, eps)
        ctx.save_for_backward(xn, mean, rstd, w, b)
        ctx.groups = groups
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        xn, mean, rstd, w, b = ctx.saved_tensors
        dyn = dy.permute(0, 2, 3, 1)
        if dyn.dtype != xn.dtype:
            dyn = dyn.to(xn.dtype)
        if not dyn.is_contiguous():
            dyn = dyn.contiguous()
        dx, dg, db = ops.groupnorm_silu_bwd(dyn, xn, mean, rstd, w, b, ctx.groups)
        return dx.permute(0, 3, 1, 2), dg, db, None, None


class _Conv3x3Fn(torch.autograd.Function):
    """3x3 convolution (padding 1, no bias) of a channels-last CUDA tensor on the library's implicit-GEMM engine (tcgen05
    in bf16: the A operand is read through 4-D TMA boxes whose out-of-bounds zero fill is the padding) -- replaces
    nn.Conv2d -> cuDNN for the heads' conv stacks (code/models/heads.py:16-42, 404-428)."""

    @staticmethod
    def forward(ctx, x, weight):
        from . import ops
        xn = x.permute(0, 2, 3, 1)                       # NHWC view of a channels-last tensor
        if not xn.is_contiguous():
            xn = xn.contiguous()
        wf, wd = ops.conv3x3_repack(weight.detach().float().contiguous(), xn.dtype)
        y = ops.conv3x3_fwd(xn, wf)
        ctx.save_for_backward(xn, wd)
        ctx.wshape = weight.shape
        ctx.wdtype = weight.dtype
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        xn, wd = ctx.saved_tensors
        dyn = dy.permute(0, 2, 3, 1)
        if dyn.dtype != xn.dtype:
            dyn = dyn.to(xn.dtype)
        if not dyn.is_contiguous():
            dyn = dyn.contiguous()
        dx = ops.conv3x3_dgrad(dyn, wd).permute(0, 3, 1, 2) if ctx.needs_input_grad[0] else None
        dw = ops.conv3x3_wgrad(dyn, xn).to(ctx.wdtype) if ctx.needs_input_grad[1] else None
        return dx, dw


def _native_conv_ok(x, conv):
    if os.environ.get("MTUS_HEAD_CONV", "native") != "native":      # A/B switch for measurements only (cuDNN path)
        return False
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float32) and conv.kernel_size == (3, 3)
            and conv.padding == (1, 1) and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
            and conv.bias is None and conv.in_channels % 64 == 0 and conv.out_channels % 64 == 0)


def _conv3x3(x, conv):
    """conv(x) on the native engine when the geometry allows it (the shipped heads always do), else the module itself."""
    if _native_conv_ok(x, conv):
        if torch.is_autocast_enabled() and x.dtype == torch.float32:
            x = x.to(torch.get_autocast_gpu_dtype())
        return _Conv3x3Fn.apply(x.contiguous(memory_format=torch.channels_last), conv.weight)
    return conv(x)


class _PointwiseConvFn(torch.autograd.Function):
    """1x1 convolution with a handful of output channels (the Conv2d(128 -> num_classes, 1) tails of the segmentation and
    detection heads) as streaming kernels over the channels-last rows: fp32 NCHW logits out; backward = one pass that writes
    dx and accumulates dw / dbias (replaces three degenerate cuBLAS GEMMs and a two-block bias reduction)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        from . import ops
        xn = x.permute(0, 2, 3, 1)
        if not xn.is_contiguous():
            xn = xn.contiguous()
        w2 = weight.detach().float().reshape(weight.shape[0], -1).contiguous()
        y = ops.pointwise_conv_fwd(xn, w2, None if bias is None else bias.detach().float().contiguous())
        ctx.save_for_backward(xn, w2)
        ctx.wshape, ctx.wdtype = weight.shape, weight.dtype
        ctx.bdtype = None if bias is None else bias.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        xn, w2 = ctx.saved_tensors
        dx, dw, db = ops.pointwise_conv_bwd(dy.float().contiguous(), xn, w2, need_dx=ctx.needs_input_grad[0])
        return (None if dx is None else dx.permute(0, 3, 1, 2), dw.reshape(ctx.wshape).to(ctx.wdtype),
                None if ctx.bdtype is None else db.to(ctx.bdtype))


def _pointwise_ok(x, conv):
    if os.environ.get("MTUS_HEAD_CONV", "native") != "native" or os.environ.get("MTUS_HEAD_POINTWISE", "1") == "0":
        return False
    l = conv.in_channels // 8
    return (isinstance(conv, nn.Conv2d) and x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float32)
            and conv.kernel_size == (1, 1) and conv.padding == (0, 0) and conv.stride == (1, 1) and conv.dilation == (1, 1)
            and conv.groups == 1 and conv.in_channels % 8 == 0 and 1 <= l <= 32 and (l & (l - 1)) == 0 and conv.out_channels <= 8)


def _conv1x1(x, conv):
    """conv(x) through the pointwise kernels when the geometry allows it (fp32 logits), else the module itself."""
    if _pointwise_ok(x, conv):
        if torch.is_autocast_enabled() and x.dtype == torch.float32:
            x = x.to(torch.get_autocast_gpu_dtype())
        return _PointwiseConvFn.apply(x, conv.weight, conv.bias)
    return conv(x)


class _BatchNormReLUFn(torch.autograd.Function):
    """relu(batch_norm(x)) for a channels-last CUDA tensor through the library's normalisation kernels: per-channel batch
    statistics (training) or the running statistics (eval), one fused normalise + ReLU pass, two-kernel backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps):
        from . import ops
        xn = x.permute(0, 2, 3, 1)
        if not xn.is_contiguous():
            xn = xn.contiguous()
        w, b = weight.float().contiguous(), bias.float().contiguous()
        if training:
            mean, rstd = ops.batchnorm_stats(xn, eps)
            with torch.no_grad():
                n = xn.numel() // xn.shape[-1]
                var = rstd.pow(-2) - eps
                running_mean.mul_(1.0 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                running_var.mul_(1.0 - momentum).add_((var * (n / max(n - 1, 1))).to(running_var.dtype), alpha=momentum)
        else:
            mean = running_mean.float().contiguous()
            rstd = (running_var.float() + eps).rsqrt().contiguous()
        y = ops.batchnorm_relu_fwd(xn, mean, rstd, w, b)
        ctx.save_for_backward(xn, y, mean, rstd, w)
        ctx.training = bool(training)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        xn, y, mean, rstd, w = ctx.saved_tensors
        dyn = dy.permute(0, 2, 3, 1)
        if dyn.dtype != xn.dtype:
            dyn = dyn.to(xn.dtype)
        if not dyn.is_contiguous():
            dyn = dyn.contiguous()
        dx, dg, db = ops.batchnorm_relu_bwd(dyn, xn, y, mean, rstd, w, training=ctx.training)
        return dx.permute(0, 3, 1, 2), dg, db, None, None, None, None, None


def _native_bn_ok(x, bn):
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float32) and x.shape[1] % 8 == 0
            and x.shape[1] // 8 <= 256 and bn.affine and bn.track_running_stats and bn.momentum is not None)


def _bn_relu(x, bn):
    if _native_bn_ok(x, bn):
        if bn.training:
            bn.num_batches_tracked.add_(1)
        return _BatchNormReLUFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.training, bn.momentum, bn.eps)
    return torch.relu(bn(x))


def _fused_gn_silu_ok(x, gn):
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float32) and x.shape[1] % 8 == 0
            and x.shape[1] // 8 <= 256 and gn.affine and x.shape[0] <= 65535)


class SegmentationHead(nn.Module):
    def __init__(self, in_channels, num_classes, kernel_size=1, upsampling=4, mid_channels=None, num_layers=2):
        super().__init__()
        mid_channels = mid_channels or in_channels
        layers, cur = [], in_channels
        for _ in range(num_layers):
            layers += [nn.Conv2d(cur, mid_channels, 3, padding=1, bias=False),
                       nn.GroupNorm(_gn_groups(mid_channels), mid_channels), nn.SiLU(inplace=True)]
            cur = mid_channels
        self.pre_head = nn.Sequential(*layers) if layers else nn.Identity()
        self.head = _SmpSegmentationHead(cur, num_classes, kernel_size=kernel_size, upsampling=upsampling)

    def forward(self, x):
        # Conv -> GroupNorm -> SiLU triples: the norm + activation run as one fused library op on CUDA (same parameters,
        # same state-dict keys); anything else (CPU, exotic shapes) takes the plain module path
        if isinstance(self.pre_head, nn.Sequential) and x.is_cuda:
            mods = list(self.pre_head)
            i = 0
            while i < len(mods):
                if (i + 2 < len(mods) and isinstance(mods[i], nn.Conv2d) and isinstance(mods[i + 1], nn.GroupNorm)
                        and isinstance(mods[i + 2], nn.SiLU)):
                    x = _conv3x3(x, mods[i])
                    gn = mods[i + 1]
                    if _fused_gn_silu_ok(x, gn):
                        x = _GroupNormSiLUFn.apply(x, gn.weight, gn.bias, gn.num_groups, gn.eps)
                    else:
                        x = mods[i + 2](gn(x))
                    i += 3
                else:
                    x = mods[i](x)
                    i += 1
            return self._tail(x)
        return self.head(self.pre_head(x))

    def _tail(self, x):
        mods = list(self.head)                   # Conv2d -> upsampling -> Identity
        if mods and isinstance(mods[0], nn.Conv2d) and _pointwise_ok(x, mods[0]):
            x = _conv1x1(x, mods[0])
            for m in mods[1:]:
                x = m(x)
            return x
        return self.head(x)


class BaselineSmpClassificationHead(nn.Module):
    def __init__(self, in_channels, num_classes, dropout=0.2):
        super().__init__()
        self.head = _SmpClassificationHead(in_channels, num_classes, dropout=dropout)

    def forward(self, features):
        return self.head(features[-1] if isinstance(features, (list, tuple)) else features)


class BaselineRegressionHead(nn.Module):
    def __init__(self, in_channels, num_points):
        super().__init__()
        self.pooling = nn.AdaptiveAvgPool2d(1)
        self.flatten = nn.Flatten()
        self.linear = nn.Linear(in_channels, num_points * 2)

    def forward(self, features):
        x = features[-1] if isinstance(features, (list, tuple)) else features
        return self.linear(self.flatten(self.pooling(x)))


class BaselineFPNGridDetectionHead(nn.Module):
    def __init__(self, fpn_out_channels, num_classes=1, num_anchors=1, mid_channels=128):
        super().__init__()
        n_out = num_anchors * (4 + num_classes)
        self.conv_block = nn.Sequential(
            nn.Conv2d(fpn_out_channels, mid_channels, 3, padding=1, bias=False), nn.BatchNorm2d(mid_channels), nn.ReLU(),
            nn.Conv2d(mid_channels, mid_channels, 3, padding=1, bias=False), nn.BatchNorm2d(mid_channels), nn.ReLU(),
            nn.Conv2d(mid_channels, n_out, 1))

    def forward(self, fpn_features):
        x = fpn_features
        if x.is_cuda:
            # Conv3x3 -> BatchNorm2d -> ReLU pairs on the native engines (same modules, same state-dict keys)
            mods = list(self.conv_block)
            i = 0
            while i < len(mods):
                if (i + 2 < len(mods) and isinstance(mods[i], nn.Conv2d) and isinstance(mods[i + 1], nn.BatchNorm2d)
                        and isinstance(mods[i + 2], nn.ReLU)):
                    x = _bn_relu(_conv3x3(x, mods[i]), mods[i + 1])
                    i += 3
                elif isinstance(mods[i], nn.Conv2d) and _pointwise_ok(x, mods[i]):
                    x = _conv1x1(x, mods[i])
                    i += 1
                else:
                    x = mods[i](x)
                    i += 1
            y = x
        else:
            y = self.conv_block(x)
        return torch.cat([torch.sigmoid(y[:, :4]), y[:, 4:]], dim=1)   # sigmoid on the 4 box channels


def build_task_head(task_config, fpn_out_channels, encoder_channels, model_config):
    name, n = task_config["task_name"], task_config["num_classes"]
    heads_cfg = (model_config or {}).get("heads", {}) or {}
    if name == "segmentation":
        c = heads_cfg.get("segmentation", {})
        if c.get("use_deep_supervision", False) or c.get("type") == "unet_like":
            raise NotImplementedError("mtus_b200 ships only the swin_b.yaml head family (SURVEY §2 row 6)")
        mid = c.get("mid_channels")
        return SegmentationHead(fpn_out_channels, n, upsampling=int(c.get("upsampling", 4)),
                                mid_channels=int(mid) if mid is not None else None, num_layers=int(c.get("num_layers", 2)))
    if name == "classification":
        c = heads_cfg.get("classification", {})
        if c.get("type") != "baseline" and not heads_cfg.get("use_baseline", False):
            raise NotImplementedError("mtus_b200 ships only the baseline classification head")
        return BaselineSmpClassificationHead(encoder_channels[-1], n, dropout=float(c.get("dropout", 0.2)))
    if name == "detection":
        c = heads_cfg.get("detection", {})
        if c.get("type", "centernet") != "baseline" and not heads_cfg.get("use_baseline", False):
            raise NotImplementedError("mtus_b200 ships only the baseline detection head")
        return BaselineFPNGridDetectionHead(fpn_out_channels, num_classes=n, mid_channels=int(c.get("mid_channels", 128)),
                                            num_anchors=int(c.get("num_anchors", 1)))
    if name == "Regression":
        c = heads_cfg.get("regression", {})
        if c.get("type") != "baseline" and not heads_cfg.get("use_baseline", False):
            raise NotImplementedError("mtus_b200 ships only the baseline regression head")
        return BaselineRegressionHead(encoder_channels[-1], n)
    raise ValueError(f"Unknown task type: {name}")


def build_all_heads(task_configs, fpn_out_channels, encoder_channels, model_config):
    heads = nn.ModuleDict()
    for cfg in task_configs:
        heads[cfg["task_id"]] = build_task_head(cfg, fpn_out_channels, encoder_channels, model_config)
    return heads
