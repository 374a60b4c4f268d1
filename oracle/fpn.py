"""Oracle: fp32 PyTorch restatement of smp's ``FPNDecoder``.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows the public algorithm of
``segmentation_models_pytorch/decoders/fpn/decoder.py`` (smp >= 0.5; absent from
/root/reference; imported at ``code/models/decoders.py:6`` and constructed at
``code/models/decoders.py:42-49``; called with a *list* of features at
``code/models/multitask_model.py:211``).  Module names reproduce smp's state-dict
keys: ``p5``, ``p{4,3,2}.skip_conv``, ``seg_blocks.{i}.block.{k}.block.{0,1}``.

Parity unpinned against smp itself (not importable here); see oracle/__init__.py.
"""

from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


class Conv3x3GNReLU(nn.Module):
    """smp ``Conv3x3GNReLU``: Conv3x3(pad 1, no bias) -> GroupNorm(32) -> ReLU [-> bilinear x2, align_corners=True]."""

    def __init__(self, in_channels: int, out_channels: int, upsample: bool = False):
        super().__init__()
        self.upsample = upsample
        self.block = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, (3, 3), stride=1, padding=1, bias=False),
            nn.GroupNorm(32, out_channels),
            nn.ReLU(inplace=True),
        )

    def forward(self, x):
        x = self.block(x)
        if self.upsample:
            x = F.interpolate(x, scale_factor=2.0, mode="bilinear", align_corners=True)
        return x


class FPNBlock(nn.Module):
    """smp ``FPNBlock``: nearest x2 of the coarser level + Conv1x1(skip) (bias on)."""

    def __init__(self, pyramid_channels: int, skip_channels: int):
        super().__init__()
        self.skip_conv = nn.Conv2d(skip_channels, pyramid_channels, kernel_size=1)

    def forward(self, x, skip):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return x + self.skip_conv(skip)


class SegmentationBlock(nn.Module):
    """smp ``SegmentationBlock``: max(1, n_upsamples) Conv3x3GNReLU blocks, each upsampling iff n_upsamples > 0."""

    def __init__(self, in_channels: int, out_channels: int, n_upsamples: int = 0):
        super().__init__()
        blocks = [Conv3x3GNReLU(in_channels, out_channels, upsample=bool(n_upsamples))]
        for _ in range(1, n_upsamples):
            blocks.append(Conv3x3GNReLU(out_channels, out_channels, upsample=True))
        self.block = nn.Sequential(*blocks)

    def forward(self, x):
        return self.block(x)


class MergeBlock(nn.Module):
    """smp ``MergeBlock``: 'add' -> elementwise sum, 'cat' -> channel concat."""

    def __init__(self, policy: str):
        super().__init__()
        if policy not in ("add", "cat"):
            raise ValueError("`merge_policy` must be one of: ['add', 'cat'], got {}".format(policy))
        self.policy = policy

    def forward(self, x: List[torch.Tensor]):
        if self.policy == "add":
            return sum(x)
        return torch.cat(x, dim=1)


class FPNDecoder(nn.Module):
    """smp ``FPNDecoder`` (ctor kwargs exactly as used at code/models/decoders.py:42-49)."""

    def __init__(self, encoder_channels, encoder_depth=5, pyramid_channels=256, segmentation_channels=128,
                 dropout=0.2, merge_policy="add"):
        super().__init__()
        self.out_channels = segmentation_channels if merge_policy == "add" else segmentation_channels * 4
        if encoder_depth < 3:
            raise ValueError("Encoder depth for FPN decoder cannot be less than 3, got {}.".format(encoder_depth))
        encoder_channels = list(encoder_channels)[::-1][: encoder_depth + 1]
        self.p5 = nn.Conv2d(encoder_channels[0], pyramid_channels, kernel_size=1)
        self.p4 = FPNBlock(pyramid_channels, encoder_channels[1])
        self.p3 = FPNBlock(pyramid_channels, encoder_channels[2])
        self.p2 = FPNBlock(pyramid_channels, encoder_channels[3])
        self.seg_blocks = nn.ModuleList([
            SegmentationBlock(pyramid_channels, segmentation_channels, n_upsamples=n) for n in [3, 2, 1, 0]])
        self.merge = MergeBlock(merge_policy)
        self.dropout = nn.Dropout2d(p=dropout, inplace=True)

    def forward(self, features: List[torch.Tensor]) -> torch.Tensor:
        c2, c3, c4, c5 = features[-4:]
        p5 = self.p5(c5)
        p4 = self.p4(p5, c4)
        p3 = self.p3(p4, c3)
        p2 = self.p2(p3, c2)
        pyramid = [blk(p) for blk, p in zip(self.seg_blocks, [p5, p4, p3, p2])]
        return self.dropout(self.merge(pyramid))
