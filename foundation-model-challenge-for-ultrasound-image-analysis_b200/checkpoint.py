"""Checkpoint / pretrained-weight interchange for the Swin encoder.

The reference's shipped configuration is ``model.encoder.pretrained: imagenet``
(``/root/reference/code/configs/swin_b.yaml:41``; the flag is computed at ``code/models/encoders.py:683`` and handed to
``timm.create_model(..., pretrained=True)`` at ``:53-59``), its checkpoints are plain ``model.state_dict()`` files
(``code/train.py:695``, reloaded at ``:735``) or dicts with a ``model_state_dict`` entry (``:714-727``).  There is no
network here, so "pretrained" means a LOCAL file:

* ``resolve_pretrained_path(name, spec)`` -- where ``pretrained: imagenet`` looks for ``<timm model name>.{pth,pt,bin,
  safetensors}`` (directories: ``$MTUS_PRETRAINED_DIR``, ``~/.cache/mtus_b200``, ``./pretrained``); ``pretrained:
  /some/file.pth`` names the file directly.
* ``convert_swin_state_dict(sd, core)`` -- maps a timm ``SwinTransformer`` state dict (current layout, the pre-0.9
  layout whose PatchMerging sits at the END of stage i, ``features_only`` ``layers_{i}`` keys, the reference's
  ``encoder.model.`` prefix) or a torchvision ``SwinTransformer`` state dict onto the native module's keys; drops the
  classifier / final norm / index and mask buffers; resizes ``relative_position_bias_table`` bicubically when the
  checkpoint's window differs (timm ``checkpoint_filter_fn`` behaviour: 224/window-7 weights into a 384/window-12 or a
  clipped-window model).
* ``load_pretrained(core, path)`` / ``load_checkpoint(model, path)``.
"""

import os
import re
from typing import Dict, Optional

import torch
import torch.nn.functional as F

_EXTS = (".pth", ".pt", ".bin", ".safetensors")


def resolve_pretrained_path(model_name: str, spec=None) -> Optional[str]:
    """File holding the weights for ``model_name``; ``spec`` is the config value of ``model.encoder.pretrained``."""
    if isinstance(spec, str) and spec not in ("imagenet", "true", "True") and os.path.isfile(spec):
        return spec
    dirs = [os.environ.get("MTUS_PRETRAINED_DIR"), os.path.expanduser("~/.cache/mtus_b200"), os.path.join(os.getcwd(), "pretrained")]
    if isinstance(spec, str) and os.path.isdir(spec):
        dirs.insert(0, spec)
    for d in dirs:
        if not d:
            continue
        for ext in _EXTS:
            p = os.path.join(d, model_name + ext)
            if os.path.isfile(p):
                return p
    return None


def read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        try:
            from safetensors.torch import load_file
        except ImportError as e:                       # pragma: no cover - depends on the image
            raise RuntimeError("mtus_b200: reading .safetensors needs the safetensors package") from e
        return load_file(path)
    obj = torch.load(path, map_location="cpu", weights_only=False)
    for key in ("model_state_dict", "state_dict", "model"):           # train.py:716 / timm / Swin official releases
        if isinstance(obj, dict) and key in obj and isinstance(obj[key], dict):
            obj = obj[key]
            break
    return obj


def resize_rel_pos_bias_table(table: torch.Tensor, new_window) -> torch.Tensor:
    """[(2w-1)^2, heads] -> [(2w'-1)^2, heads] by bicubic interpolation of the (2w-1) x (2w-1) offset grid
    (timm ``resize_rel_pos_bias_table_simple`` / the Swin fine-tuning recipe)."""
    n_old, heads = table.shape
    s_old = int(round(n_old ** 0.5))
    if s_old * s_old != n_old:
        raise ValueError("relative_position_bias_table is not a square grid")
    new_h, new_w = (new_window, new_window) if isinstance(new_window, int) else new_window
    sh, sw = 2 * new_h - 1, 2 * new_w - 1
    if (sh, sw) == (s_old, s_old):
        return table
    grid = table.float().permute(1, 0).reshape(1, heads, s_old, s_old)
    out = F.interpolate(grid, size=(sh, sw), mode="bicubic", align_corners=False)
    return out.reshape(heads, sh * sw).permute(1, 0).contiguous().to(table.dtype)


_TV_BLOCK = re.compile(r"^features\.(\d+)\.(\d+)\.(.*)$")


def _from_torchvision(sd):
    """torchvision.models.swin_transformer key layout -> timm key layout (SURVEY §8c key map)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("features.0.0."):
            out["patch_embed.proj." + k[len("features.0.0."):]] = v
            continue
        if k.startswith("features.0.2."):
            out["patch_embed.norm." + k[len("features.0.2."):]] = v
            continue
        mt2 = re.match(r"^features\.(2|4|6)\.(norm|reduction)\.(.*)$", k)
        if mt2:
            out[f"layers.{int(mt2.group(1)) // 2}.downsample.{mt2.group(2)}.{mt2.group(3)}"] = v
            continue
        mt = _TV_BLOCK.match(k)
        if mt and int(mt.group(1)) in (1, 3, 5, 7):
            rest = mt.group(3).replace("mlp.0.", "mlp.fc1.").replace("mlp.3.", "mlp.fc2.")
            out[f"layers.{(int(mt.group(1)) - 1) // 2}.blocks.{mt.group(2)}.{rest}"] = v
    return out                                         # norm.*, head.* are dropped


def convert_swin_state_dict(sd: Dict[str, torch.Tensor], core) -> Dict[str, torch.Tensor]:
    """Returns a state dict with exactly the keys of ``core`` (a ``SwinCore``); raises on missing / mis-shaped tensors."""
    sd = dict(sd)
    if any(k.startswith("features.0.0.") for k in sd):
        sd = _from_torchvision(sd)
    norm = {}
    for k, v in sd.items():
        for prefix in ("encoder.model.", "model.", "module.", "backbone."):
            if k.startswith(prefix):
                k = k[len(prefix):]
        k = re.sub(r"^layers_(\d+)\.", r"layers.\1.", k)
        norm[k] = v
    old_layout = any(k.startswith("layers.0.downsample.") for k in norm)   # timm < 0.9: merge at the end of stage i
    want = core.state_dict()
    out = {}
    for key, ref in want.items():
        src = re.sub(r"^layers_(\d+)\.", r"layers.\1.", key)
        if old_layout and ".downsample." in src:
            i = int(src.split(".")[1])
            src = src.replace(f"layers.{i}.", f"layers.{i - 1}.", 1)
        if src not in norm:
            raise KeyError(f"mtus_b200: checkpoint has no tensor for {key!r} (looked for {src!r})")
        v = norm[src]
        if key.endswith("relative_position_bias_table") and v.shape != ref.shape:
            side = int(round(ref.shape[0] ** 0.5))
            v = resize_rel_pos_bias_table(v, (side + 1) // 2)
        if tuple(v.shape) != tuple(ref.shape):
            raise ValueError(f"mtus_b200: {key}: checkpoint shape {tuple(v.shape)} != model shape {tuple(ref.shape)}")
        out[key] = v.to(ref.dtype)
    return out


def load_pretrained(core, path: str):
    """Loads a timm / torchvision / reference checkpoint file into a ``SwinCore`` (in place)."""
    core.load_state_dict(convert_swin_state_dict(read_state_dict(path), core))
    return core


def load_checkpoint(model, path: str, strict: bool = True):
    """``model.load_state_dict(torch.load(best_model.pth))`` (train.py:735) also accepting the periodic checkpoint dicts
    (train.py:714-727)."""
    return model.load_state_dict(read_state_dict(path), strict=strict)
