"""Oracle: restatement of the smp pieces the reference's heads / losses import.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  These sit DOWNSTREAM of the hot
path (out of scope) and exist only so that the reference's
own ``code/models/heads.py`` (imports at :6, uses at :33,85,94,134,366) and
``code/losses/loss_functions.py`` (:7,169,173) can run on top of the shims, and so
the 27-task CPU baseline can be timed.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F


class SegmentationHead(nn.Sequential):
    """smp ``base.SegmentationHead``: Conv2d(k, pad k//2) -> UpsamplingBilinear2d(scale) -> Identity."""

    def __init__(self, in_channels, out_channels, kernel_size=3, activation=None, upsampling=1):
        conv2d = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=kernel_size // 2)
        up = nn.UpsamplingBilinear2d(scale_factor=upsampling) if upsampling > 1 else nn.Identity()
        if activation is not None:
            raise NotImplementedError("oracle: only activation=None is used by the reference")
        super().__init__(conv2d, up, nn.Identity())


class ClassificationHead(nn.Sequential):
    """smp ``base.ClassificationHead``: pool -> Flatten -> Dropout -> Linear -> Identity."""

    def __init__(self, in_channels, classes, pooling="avg", dropout=0.2, activation=None):
        if pooling not in ("max", "avg"):
            raise ValueError("Pooling should be one of ('max', 'avg'), got {}.".format(pooling))
        pool = nn.AdaptiveAvgPool2d(1) if pooling == "avg" else nn.AdaptiveMaxPool2d(1)
        drop = nn.Dropout(p=dropout, inplace=True) if dropout else nn.Identity()
        if activation is not None:
            raise NotImplementedError("oracle: only activation=None is used by the reference")
        super().__init__(pool, nn.Flatten(), drop, nn.Linear(in_channels, classes, bias=True), nn.Identity())


class DiceLoss(nn.Module):
    """smp ``losses.DiceLoss`` for the modes the reference uses (multiclass; binary kept for completeness).

    multiclass: probs = log_softmax(dim=1).exp(); one-hot target; sums over dims (0, 2);
    smooth=0, eps=1e-7; loss = 1 - dice; classes absent from the target are masked; mean over classes.
    """

    def __init__(self, mode="multiclass", smooth=0.0, eps=1e-7, from_logits=True):
        super().__init__()
        assert mode in ("binary", "multiclass")
        self.mode, self.smooth, self.eps, self.from_logits = mode, smooth, eps, from_logits

    def forward(self, y_pred, y_true):
        assert y_true.size(0) == y_pred.size(0)
        if self.from_logits:
            y_pred = y_pred.log_softmax(dim=1).exp() if self.mode == "multiclass" else F.logsigmoid(y_pred).exp()
        bs, num_classes, dims = y_true.size(0), y_pred.size(1), (0, 2)
        if self.mode == "binary":
            y_true = y_true.view(bs, 1, -1)
            y_pred = y_pred.view(bs, 1, -1)
        else:
            y_true = F.one_hot(y_true.view(bs, -1).long(), num_classes).permute(0, 2, 1)
            y_pred = y_pred.view(bs, num_classes, -1)
        y_true = y_true.type_as(y_pred)
        inter = torch.sum(y_pred * y_true, dim=dims)
        card = torch.sum(y_pred + y_true, dim=dims)
        score = (2.0 * inter + self.smooth) / (card + self.smooth).clamp_min(self.eps)
        loss = 1.0 - score
        loss = loss * (y_true.sum(dims) > 0).to(loss.dtype)
        return loss.mean()
