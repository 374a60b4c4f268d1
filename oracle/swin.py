"""Oracle: fp32 PyTorch restatement of timm's Swin Transformer (features_only).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows the public algorithm of
``timm.models.swin_transformer`` (timm >= 1.0; absent from /root/reference,
called at ``code/models/encoders.py:53-59`` and ``:104``).  Each class names
the timm symbol it restates.  Module/parameter names reproduce timm's
``FeatureListNet(flatten_sequential=True)`` state-dict keys
(``patch_embed.*``, ``layers_{i}.downsample.*``, ``layers_{i}.blocks.{j}.*``)
because the reference checkpoints with ``model.state_dict()``
(``code/train.py:695``).
"""

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# timm names -> (embed_dim, depths, heads, window).  ``code/models/encoders.py:14-19``
# maps swin_t/s/b/l onto the *_window7_224 names; any other ``swin_*`` string is
# passed through verbatim (encoders.py:51), e.g. swin_large_patch4_window12_384.
SWIN_VARIANTS = {
    "swin_tiny_patch4_window7_224": (96, (2, 2, 6, 2), (3, 6, 12, 24), 7),
    "swin_small_patch4_window7_224": (96, (2, 2, 18, 2), (3, 6, 12, 24), 7),
    "swin_base_patch4_window7_224": (128, (2, 2, 18, 2), (4, 8, 16, 32), 7),
    "swin_large_patch4_window7_224": (192, (2, 2, 18, 2), (6, 12, 24, 48), 7),
    "swin_base_patch4_window12_384": (128, (2, 2, 18, 2), (4, 8, 16, 32), 12),
    "swin_large_patch4_window12_384": (192, (2, 2, 18, 2), (6, 12, 24, 48), 12),
    # not a timm name: a micro variant (head_dim 32 like every Swin) used only by the parity tests /
    # golden fixtures so that they stay small; exercises odd-size merging and window clipping.
    "swin_micro_patch4_window7_test": (32, (2, 2, 2, 2), (1, 2, 4, 8), 7),
}


def window_partition(x: torch.Tensor, window: Tuple[int, int]) -> torch.Tensor:
    """timm ``window_partition``: [B,H,W,C] -> [B*nW, wh, ww, C]."""
    B, H, W, C = x.shape
    x = x.view(B, H // window[0], window[0], W // window[1], window[1], C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window[0], window[1], C)


def window_reverse(windows: torch.Tensor, window: Tuple[int, int], H: int, W: int) -> torch.Tensor:
    """timm ``window_reverse``: [B*nW, wh, ww, C] -> [B,H,W,C]."""
    C = windows.shape[-1]
    x = windows.view(-1, H // window[0], W // window[1], window[0], window[1], C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, H, W, C)


def relative_position_index(win_h: int, win_w: int) -> torch.Tensor:
    """timm ``get_relative_position_index``: idx[i,j]=(yi-yj+wh-1)*(2ww-1)+(xi-xj+ww-1)."""
    coords = torch.stack(torch.meshgrid(torch.arange(win_h), torch.arange(win_w), indexing="ij"))
    flat = coords.flatten(1)
    rel = flat[:, :, None] - flat[:, None, :]
    rel = rel.permute(1, 2, 0).contiguous()
    rel[:, :, 0] += win_h - 1
    rel[:, :, 1] += win_w - 1
    rel[:, :, 0] *= 2 * win_w - 1
    return rel.sum(-1)


class Mlp(nn.Module):
    """timm ``Mlp``: fc1 -> exact-erf GELU -> fc2 (dropout p=0)."""

    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class WindowAttention(nn.Module):
    """timm ``WindowAttention`` (non-fused branch, the default)."""

    def __init__(self, dim: int, num_heads: int, window: Tuple[int, int]):
        super().__init__()
        self.dim = dim
        self.window_size = window
        self.window_area = window[0] * window[1]
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window[0] - 1) * (2 * window[1] - 1), num_heads))
        self.register_buffer("relative_position_index",
                             relative_position_index(window[0], window[1]), persistent=False)
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    def _get_rel_pos_bias(self) -> torch.Tensor:
        bias = self.relative_position_bias_table[self.relative_position_index.view(-1)]
        bias = bias.view(self.window_area, self.window_area, -1).permute(2, 0, 1).contiguous()
        return bias.unsqueeze(0)

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B_, N, C = x.shape
        qkv = self.qkv(x).reshape(B_, N, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q = q * self.scale
        attn = q @ k.transpose(-2, -1)
        attn = attn + self._get_rel_pos_bias()
        if mask is not None:
            nW = mask.shape[0]
            attn = attn.view(-1, nW, self.num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, self.num_heads, N, N)
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B_, N, -1)
        return self.proj(x)


def drop_path(x: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """timm ``drop_path``: per-sample Bernoulli(keep) / keep, identity in eval."""
    if p == 0.0 or not training:
        return x
    keep = 1.0 - p
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    mask = x.new_empty(shape).bernoulli_(keep)
    if keep > 0.0:
        mask.div_(keep)
    return x * mask


class SwinTransformerBlock(nn.Module):
    """timm ``SwinTransformerBlock`` incl. ``_calc_window_shift``, ``get_attn_mask``, ``_attn``."""

    def __init__(self, dim, input_resolution, num_heads, window_size, shift_size, mlp_ratio=4.0,
                 drop_path_rate=0.0):
        super().__init__()
        self.dim = dim
        self.input_resolution = tuple(input_resolution)
        # _calc_window_shift: window clipped to the feature map; shift off when res <= window
        tw, ts = (window_size, window_size), (shift_size, shift_size)
        ws = [r if r <= w else w for r, w in zip(self.input_resolution, tw)]
        ss = [0 if r <= w else s for r, w, s in zip(self.input_resolution, ws, ts)]
        self.window_size, self.shift_size = tuple(ws), tuple(ss)
        self.window_area = self.window_size[0] * self.window_size[1]
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, num_heads, self.window_size)
        self.drop_path_rate = float(drop_path_rate)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.register_buffer("attn_mask", self.get_attn_mask(), persistent=False)

    def get_attn_mask(self) -> Optional[torch.Tensor]:
        if not any(self.shift_size):
            return None
        H, W = self.input_resolution
        H = math.ceil(H / self.window_size[0]) * self.window_size[0]
        W = math.ceil(W / self.window_size[1]) * self.window_size[1]
        img_mask = torch.zeros((1, H, W, 1))
        cnt = 0
        for h in ((0, -self.window_size[0]), (-self.window_size[0], -self.shift_size[0]),
                  (-self.shift_size[0], None)):
            for w in ((0, -self.window_size[1]), (-self.window_size[1], -self.shift_size[1]),
                      (-self.shift_size[1], None)):
                img_mask[:, h[0]:h[1], w[0]:w[1], :] = cnt
                cnt += 1
        mw = window_partition(img_mask, self.window_size).view(-1, self.window_area)
        am = mw.unsqueeze(1) - mw.unsqueeze(2)
        return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)

    def _attn(self, x: torch.Tensor) -> torch.Tensor:
        B, H, W, C = x.shape
        has_shift = any(self.shift_size)
        # roll FIRST, then pad (timm order; torchvision/HF pad first -- see SURVEY A11)
        if has_shift:
            x = torch.roll(x, shifts=(-self.shift_size[0], -self.shift_size[1]), dims=(1, 2))
        pad_h = (self.window_size[0] - H % self.window_size[0]) % self.window_size[0]
        pad_w = (self.window_size[1] - W % self.window_size[1]) % self.window_size[1]
        x = F.pad(x, (0, 0, 0, pad_w, 0, pad_h))
        Hp, Wp = x.shape[1], x.shape[2]
        xw = window_partition(x, self.window_size).view(-1, self.window_area, C)
        aw = self.attn(xw, mask=self.attn_mask)
        aw = aw.view(-1, self.window_size[0], self.window_size[1], C)
        x = window_reverse(aw, self.window_size, Hp, Wp)[:, :H, :W, :].contiguous()
        if has_shift:
            x = torch.roll(x, shifts=self.shift_size, dims=(1, 2))
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, H, W, C = x.shape
        x = x + drop_path(self._attn(self.norm1(x)), self.drop_path_rate, self.training)
        x = x.reshape(B, -1, C)
        x = x + drop_path(self.mlp(self.norm2(x)), self.drop_path_rate, self.training)
        return x.reshape(B, H, W, C)


class PatchMerging(nn.Module):
    """timm ``PatchMerging``: pad-to-even, 2x2 gather (order h0w0,h1w0,h0w1,h1w1), LN(4C), Linear(4C->2C, no bias)."""

    def __init__(self, dim: int, out_dim: int):
        super().__init__()
        self.norm = nn.LayerNorm(4 * dim)
        self.reduction = nn.Linear(4 * dim, out_dim, bias=False)

    def forward(self, x):
        B, H, W, C = x.shape
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
        _, H, W, _ = x.shape
        x = x.reshape(B, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 4, 2, 5).flatten(3)
        return self.reduction(self.norm(x))


class SwinTransformerStage(nn.Module):
    """timm ``SwinTransformerStage``: downsample FIRST (stages 1..3), then the blocks."""

    def __init__(self, dim, out_dim, input_resolution, depth, downsample, num_heads, window_size,
                 drop_path: Sequence[float]):
        super().__init__()
        if downsample:
            self.downsample = PatchMerging(dim, out_dim)
            self.output_resolution = tuple((i + 1) // 2 for i in input_resolution)
        else:
            assert dim == out_dim
            self.downsample = nn.Identity()
            self.output_resolution = tuple(input_resolution)
        self.blocks = nn.Sequential(*[
            SwinTransformerBlock(out_dim, self.output_resolution, num_heads, window_size,
                                 0 if i % 2 == 0 else window_size // 2, drop_path_rate=drop_path[i])
            for i in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class PatchEmbed(nn.Module):
    """timm ``PatchEmbed`` (NHWC output, strict image size): Conv2d(3,C,4,4) -> NHWC -> LayerNorm."""

    def __init__(self, img_size: int, patch: int, in_chans: int, embed_dim: int):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.grid_size = (img_size // patch, img_size // patch)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch, stride=patch)
        self.norm = nn.LayerNorm(embed_dim)

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0], f"Input height ({H}) doesn't match model ({self.img_size[0]})."
        assert W == self.img_size[1], f"Input width ({W}) doesn't match model ({self.img_size[1]})."
        x = self.proj(x).permute(0, 2, 3, 1)
        return self.norm(x)


class _FeatureInfo:
    def __init__(self, chans, reds):
        self._c, self._r = list(chans), list(reds)

    def channels(self):
        return list(self._c)

    def reduction(self):
        return list(self._r)


class SwinFeatures(nn.Module):
    """timm ``SwinTransformer`` wrapped by ``FeatureListNet(out_indices=(0,1,2,3))``.

    Returns the four stage outputs (NHWC, strides 4/8/16/32) WITHOUT the model's
    final ``norm``.  Default ``drop_path_rate=0.1`` (timm default; the
    reference passes no drop kwargs, encoders.py:53-59).
    """

    def __init__(self, name: str, img_size: int = 224, drop_path_rate: float = 0.1,
                 out_indices=(0, 1, 2, 3)):
        super().__init__()
        if name not in SWIN_VARIANTS:
            raise RuntimeError(f"Unknown model ({name})")
        embed_dim, depths, heads, window = SWIN_VARIANTS[name]
        self.embed_dim, self.depths, self.num_heads, self.window = embed_dim, depths, heads, window
        self.out_indices = tuple(out_indices)
        self.patch_embed = PatchEmbed(img_size, 4, 3, embed_dim)
        dpr = [r.tolist() for r in torch.linspace(0, drop_path_rate, sum(depths)).split(depths)]
        res = self.patch_embed.grid_size
        in_dim = embed_dim
        chans, reds = [], []
        for i in range(4):
            out_dim = embed_dim * 2 ** i
            stage = SwinTransformerStage(in_dim, out_dim, res, depths[i], i > 0, heads[i], window, dpr[i])
            setattr(self, f"layers_{i}", stage)
            res, in_dim = stage.output_resolution, out_dim
            chans.append(out_dim)
            reds.append(4 * 2 ** i)
        self.feature_info = _FeatureInfo([chans[i] for i in self.out_indices],
                                         [reds[i] for i in self.out_indices])
        self.apply(self._init)

    @staticmethod
    def _init(m):
        # timm swin init_weights: trunc_normal_(std=.02) for Linear weights, zero bias
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    def forward(self, x) -> List[torch.Tensor]:
        x = self.patch_embed(x)
        outs = []
        for i in range(4):
            x = getattr(self, f"layers_{i}")(x)
            if i in self.out_indices:
                outs.append(x)
        return outs


def create_model(name, pretrained=False, features_only=False, out_indices=(0, 1, 2, 3), img_size=224,
                 drop_path_rate=0.1, **kwargs):
    """Shim for ``timm.create_model`` restricted to what encoders.py:53-59 uses."""
    if pretrained:
        raise RuntimeError("oracle: pretrained weights are not available offline; set model.encoder.pretrained: null")
    if not features_only:
        raise NotImplementedError("oracle restates only the features_only path used by the reference")
    return SwinFeatures(name, img_size=img_size, drop_path_rate=drop_path_rate, out_indices=out_indices)


# --------------------------------------------------------------------------------------
# cross-check helper: load oracle weights into torchvision's independent implementation
# --------------------------------------------------------------------------------------

def to_torchvision_state_dict(sd: dict) -> dict:
    """Map oracle/timm keys to ``torchvision.models.swin_transformer.SwinTransformer.features`` keys.

    tv layout: features.0 = [conv, Permute, LN]; features.{1,3,5,7} = stage blocks;
    features.{2,4,6} = PatchMerging placed AFTER a stage (same dataflow as timm's
    downsample-first, only the names differ -- SURVEY A8).
    """
    out = {}
    for k, v in sd.items():
        if k.startswith("patch_embed.proj."):
            out["0.0." + k.split(".", 2)[2]] = v
        elif k.startswith("patch_embed.norm."):
            out["0.2." + k.split(".", 2)[2]] = v
        elif k.startswith("layers_"):
            i = int(k[len("layers_")])
            rest = k.split(".", 1)[1]
            if rest.startswith("downsample."):
                out[f"{2 * i}." + rest[len("downsample."):]] = v
            else:
                _, j, tail = rest.split(".", 2)
                tail = tail.replace("mlp.fc1", "mlp.0").replace("mlp.fc2", "mlp.3")
                out[f"{2 * i + 1}.{j}.{tail}"] = v
    return out
