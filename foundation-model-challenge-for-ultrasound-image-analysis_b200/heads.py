"""Task heads used to drive the hot path (PyTorch; OUT of the kernel scope -- SURVEY §2 row 6, §8f N1).

Only the head types selected by ``configs/swin_b.yaml`` are provided (the "baseline" family of
``/root/reference/code/models/heads.py``): the standard segmentation head (heads.py:16-42, reached
because ``type: baseline`` falls through at heads.py:478-496), ``BaselineSmpClassificationHead``
(:363-385), ``BaselineRegressionHead`` (:388-407) and ``BaselineFPNGridDetectionHead`` (:410-434).
Module / parameter names match the reference so its checkpoints load.  Under the reference's own
``MultiTaskModel`` the reference's heads are used instead; these exist because /root/reference (and
smp) are not available on the benchmark box.
"""

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
                    x = mods[i](x)
                    gn = mods[i + 1]
                    if _fused_gn_silu_ok(x, gn):
                        x = _GroupNormSiLUFn.apply(x, gn.weight, gn.bias, gn.num_groups, gn.eps)
                    else:
                        x = mods[i + 2](gn(x))
                    i += 3
                else:
                    x = mods[i](x)
                    i += 1
            return self.head(x)
        return self.head(self.pre_head(x))


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
        y = self.conv_block(fpn_features)
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
