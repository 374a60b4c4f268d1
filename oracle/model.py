"""Oracle: stand-alone fp32 restatement of the reference's MultiTaskModel + one training step.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  /root/reference does not exist on
the GPU box, so the CPU baseline that bench.py times there cannot import the
reference's ``code/models``; this module restates, in its own words, the routing of
``MultiTaskModel.forward`` (code/models/multitask_model.py:176-250), the encoder
wrapper (code/models/encoders.py:103-109), the decoder factory
(code/models/decoders.py:63-103), the *baseline* heads selected by
``configs/swin_b.yaml`` (code/models/heads.py:16-42, 366-428, 478-560) and the loss /
step semantics of ``train_epoch`` (code/train.py:343-455).  In the authoring
container tests/test_reference_dropin.py checks it against the reference's own
classes running on the shims.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import swin, fpn, heads as smp_heads

SHORT_NAMES = {  # code/models/encoders.py:14-19
    "swin_t": "swin_tiny_patch4_window7_224",
    "swin_s": "swin_small_patch4_window7_224",
    "swin_b": "swin_base_patch4_window7_224",
    "swin_l": "swin_large_patch4_window7_224",
}


def _gn_groups(c):
    g = min(32, c)
    while c % g:
        g -= 1
    return g


class OracleSwinEncoder(nn.Module):
    """encoders.py:37-160 without MoE: timm features -> NCHW contiguous list."""
    is_timm_encoder = True

    def __init__(self, model_name="swin_b", img_size=224, drop_path_rate=0.1):
        super().__init__()
        self.model = swin.create_model(SHORT_NAMES.get(model_name, model_name), pretrained=False,
                                       features_only=True, out_indices=(0, 1, 2, 3), img_size=img_size,
                                       drop_path_rate=drop_path_rate)
        self._out_channels = self.model.feature_info.channels()
        self.output_stride = 32
        self.supports_task_id = False
        self.handles_moe = False
        self.use_moe = False

    def forward(self, x, task_id=None):
        return [f.permute(0, 3, 1, 2).contiguous() for f in self.model(x)]

    @property
    def out_channels(self):
        return [3] + list(self._out_channels)


class SegHead(nn.Module):
    """heads.py:16-42: (Conv3x3 no-bias -> GN -> SiLU) x num_layers -> smp SegmentationHead(k=1, up)."""

    def __init__(self, in_ch, num_classes, upsampling=4, mid=None, num_layers=2):
        super().__init__()
        mid = mid or in_ch
        seq, c = [], in_ch
        for _ in range(num_layers):
            seq += [nn.Conv2d(c, mid, 3, padding=1, bias=False), nn.GroupNorm(_gn_groups(mid), mid), nn.SiLU(inplace=True)]
            c = mid
        self.pre_head = nn.Sequential(*seq) if seq else nn.Identity()
        self.head = smp_heads.SegmentationHead(c, num_classes, kernel_size=1, upsampling=upsampling)

    def forward(self, x):
        return self.head(self.pre_head(x))


class ClsHead(nn.Module):
    """heads.py:363-385 (baseline): smp ClassificationHead on features[-1]."""

    def __init__(self, in_ch, num_classes, dropout=0.2):
        super().__init__()
        self.head = smp_heads.ClassificationHead(in_ch, num_classes, pooling="avg", dropout=dropout)

    def forward(self, feats):
        return self.head(feats[-1] if isinstance(feats, (list, tuple)) else feats)


class RegHead(nn.Module):
    """heads.py:388-407 (baseline): GAP -> Linear(in, 2*num_points)."""

    def __init__(self, in_ch, num_points):
        super().__init__()
        self.pooling, self.flatten = nn.AdaptiveAvgPool2d(1), nn.Flatten()
        self.linear = nn.Linear(in_ch, num_points * 2)

    def forward(self, feats):
        x = feats[-1] if isinstance(feats, (list, tuple)) else feats
        return self.linear(self.flatten(self.pooling(x)))


class DetHead(nn.Module):
    """heads.py:410-434 (baseline): (Conv3x3-BN-ReLU) x2 -> Conv1x1; sigmoid on the 4 box channels."""

    def __init__(self, in_ch, num_classes=1, num_anchors=1, mid=128):
        super().__init__()
        n_out = num_anchors * (4 + num_classes)
        self.conv_block = nn.Sequential(
            nn.Conv2d(in_ch, mid, 3, padding=1, bias=False), nn.BatchNorm2d(mid), nn.ReLU(),
            nn.Conv2d(mid, mid, 3, padding=1, bias=False), nn.BatchNorm2d(mid), nn.ReLU(),
            nn.Conv2d(mid, n_out, 1))

    def forward(self, x):
        y = self.conv_block(x)
        return torch.cat([torch.sigmoid(y[:, :4]), y[:, 4:]], dim=1)


def build_head(task_cfg, fpn_out, enc_channels, cfg):
    name, n = task_cfg["task_name"], task_cfg["num_classes"]
    h = cfg.get("model.heads", {}) or {}
    if name == "segmentation":
        c = h.get("segmentation", {})
        mid = c.get("mid_channels")
        return SegHead(fpn_out, n, int(c.get("upsampling", 4)), int(mid) if mid is not None else None,
                       int(c.get("num_layers", 2)))
    if name == "classification":
        return ClsHead(enc_channels[-1], n, float(h.get("classification", {}).get("dropout", 0.2)))
    if name == "detection":
        c = h.get("detection", {})
        return DetHead(fpn_out, n, int(c.get("num_anchors", 1)), int(c.get("mid_channels", 128)))
    if name == "Regression":
        return RegHead(enc_channels[-1], n)
    raise ValueError(f"Unknown task type: {name}")


class OracleTaskPrompt(nn.Module):
    """task_prompt.py:28-143, literally: one multi-hot row per task ([task types | num_classes_<n> tags | id tokens without the
    "T<k>" prefix], every vocabulary sorted), expanded to B rows -> Linear -> [B, channels, p, p] -> tanh -> bilinear
    (align_corners=False) to the image size -> x + s * prompt ("add") or x * (1 + s * prompt) ("mul")."""

    def __init__(self, task_configs, channels, prompt_size, mode, init_scale, use_tanh):
        super().__init__()
        import re
        drop = re.compile(r"^t\d+[a-z]?$", re.IGNORECASE)
        ids = [str(t["task_id"]) for t in task_configs]
        kinds = [str(t.get("task_name", "unknown")).lower() for t in task_configs]
        tags = [f"num_classes_{int(t.get('num_classes', -1))}" for t in task_configs]
        toks = [[w for w in (q.strip().lower() for q in i.split("_")) if w and not drop.match(w)] for i in ids]
        vk, vt, vw = sorted(set(kinds)), sorted(set(tags)), sorted(set(sum(toks, [])))
        meta = torch.zeros(len(ids), len(vk) + len(vt) + len(vw))
        for r in range(len(ids)):
            meta[r, vk.index(kinds[r])] = 1.0
            meta[r, len(vk) + vt.index(tags[r])] = 1.0
            for w in toks[r]:
                meta[r, len(vk) + len(vt) + vw.index(w)] = 1.0
        self.rows = {t: r for r, t in enumerate(ids)}
        self.channels, self.p, self.mode, self.use_tanh = channels, prompt_size, mode, use_tanh
        self.register_buffer("task_metadata", meta)
        self.prompt_proj = nn.Linear(meta.shape[1], channels * prompt_size * prompt_size)
        self.prompt_scale = nn.Parameter(torch.tensor(float(init_scale)))

    def inject(self, x, task_id):
        B = x.shape[0]
        vec = self.task_metadata[self.rows[task_id]].to(x.device).unsqueeze(0).expand(B, -1)
        pr = self.prompt_proj(vec).view(B, self.channels, self.p, self.p)
        if self.use_tanh:
            pr = torch.tanh(pr)
        if pr.shape[-2:] != x.shape[-2:]:
            pr = F.interpolate(pr, size=x.shape[-2:], mode="bilinear", align_corners=False)
        s, pr = self.prompt_scale.to(x.dtype), pr.to(x.dtype)
        return x + s * pr if self.mode == "add" else x * (1.0 + s * pr)


class OracleMultiTaskModel(nn.Module):
    """multitask_model.py:13-250 for the swin_b.yaml family of configs (FiLM and TaskPrompt2D optional; no MoE)."""

    def __init__(self, cfg, drop_path_rate=0.1):
        super().__init__()
        self.task_configs = cfg.get_task_configs()
        self.encoder = OracleSwinEncoder(cfg.get("model.encoder.name"), cfg.get("data.image_size", 224), drop_path_rate)
        ch = self.encoder.out_channels

        def mk():
            return fpn.FPNDecoder(ch, len(ch) - 1, int(cfg.get("model.decoder.pyramid_channels", 256)),
                                  int(cfg.get("model.decoder.segmentation_channels", 128)),
                                  float(cfg.get("model.decoder.dropout", 0.2)),
                                  cfg.get("model.decoder.merge_policy", "cat"))

        self.fpn_decoder_seg = mk()
        self.fpn_decoder_det = mk() if cfg.get("model.decoder.separate_detection_fpn", True) else self.fpn_decoder_seg
        self.fpn_decoder_cls = mk() if cfg.get("model.decoder.separate_classification_fpn", True) else self.fpn_decoder_seg
        self.fpn_decoder_reg = mk() if cfg.get("model.decoder.separate_regression_fpn", True) else self.fpn_decoder_seg
        self.use_fpn_for_cls = cfg.get("model.decoder.use_fpn_for_classification", True)
        self.use_fpn_for_reg = cfg.get("model.decoder.use_fpn_for_regression", True)
        if self.use_fpn_for_cls or self.use_fpn_for_reg:
            raise NotImplementedError("oracle model restates the swin_b.yaml routing (cls/reg bypass the FPN)")
        self.fpn_out_channels = self.fpn_decoder_seg.out_channels
        # FiLM on the decoder output (multitask_model.py:52-79; film_layer.py:103-148 per-task parameters, :151-216 embedding
        # generator): out = gamma[c] * x + beta[c]
        self.use_film = bool(cfg.get("model.use_film", False))
        if self.use_film:
            fc = cfg.get("model.film", {}) or {}
            ids = [t["task_id"] for t in self.task_configs]
            self.film_affine = bool(fc.get("use_affine", True))
            self.film_embedding = bool(fc.get("use_task_embedding", False))
            gen = nn.Module()
            C = self.fpn_out_channels
            if self.film_embedding:
                E = int(fc.get("embedding_dim", 64))
                self._film_idx = {t: i for i, t in enumerate(ids)}
                gen.task_embeddings = nn.Embedding(len(ids), E)
                gen.gamma_generator = nn.Sequential(nn.Linear(E, 2 * C), nn.ReLU(), nn.Linear(2 * C, C))
                if self.film_affine:
                    gen.beta_generator = nn.Sequential(nn.Linear(E, 2 * C), nn.ReLU(), nn.Linear(2 * C, C))
            else:
                gen.task_gammas = nn.ParameterDict({t: nn.Parameter(torch.ones(C)) for t in ids})
                if self.film_affine:
                    gen.task_betas = nn.ParameterDict({t: nn.Parameter(torch.zeros(C)) for t in ids})
            self.film_generator = gen
        # input-level task prompt (multitask_model.py:81-111; task_prompt.py): built before the heads, like the reference
        tp = cfg.get("model.task_prompt", {}) or {}
        self.use_task_prompt = bool(tp.get("enabled", False))
        names = tp.get("apply_to_task_names", None)
        self.task_prompt_apply_task_names = None if names is None else {str(n).lower() for n in names}
        if self.use_task_prompt:
            self.task_prompt = OracleTaskPrompt(self.task_configs, int(tp.get("channels", 1)), int(tp.get("prompt_size", 32)),
                                                str(tp.get("inject_mode", "add")).lower(), float(tp.get("init_scale", 0.1)),
                                                bool(tp.get("use_tanh", True)))
        self.heads = nn.ModuleDict({t["task_id"]: build_head(t, self.fpn_out_channels, ch, cfg)
                                    for t in self.task_configs})
        self.task_id_to_name = {t["task_id"]: t["task_name"] for t in self.task_configs}

    def _film(self, x, task_id):
        if not self.use_film:
            return x
        g = self.film_generator
        if self.film_embedding:
            e = g.task_embeddings.weight[self._film_idx[task_id]]
            gamma, beta = g.gamma_generator(e), (g.beta_generator(e) if self.film_affine else None)
        else:
            gamma, beta = g.task_gammas[task_id], (g.task_betas[task_id] if self.film_affine else None)
        x = gamma.view(1, -1, 1, 1) * x
        return x + beta.view(1, -1, 1, 1) if beta is not None else x

    def forward(self, x, task_id):
        if task_id not in self.heads:
            raise ValueError(f"Unknown task_id: {task_id}")
        name = self.task_id_to_name[task_id]
        if self.use_task_prompt and (self.task_prompt_apply_task_names is None
                                     or name.lower() in self.task_prompt_apply_task_names):   # multitask_model.py:194-199
            x = self.task_prompt.inject(x, task_id)
        feats = self.encoder(x)
        if name == "segmentation":
            return self.heads[task_id](self._film(self.fpn_decoder_seg(feats), task_id))
        if name == "detection":
            return self.heads[task_id](self._film(self.fpn_decoder_det(feats), task_id))
        return self.heads[task_id](feats)


# ---- losses / step (code/train.py:343-455, code/losses/loss_functions.py) ----------------------

def detection_loss(outputs, labels, cls_w=2.0, box_w=1.0):
    """train.py:395-418 + DetectionLoss (loss_functions.py:10-53): sample the grid cell under the GT
    centre, BCE on objectness + SmoothL1 on the box of positive samples."""
    B, C, H, W = outputs.shape
    cx, cy = (labels[:, 0] + labels[:, 2]) / 2.0, (labels[:, 1] + labels[:, 3]) / 2.0
    ih = torch.clamp((cy * H).long(), 0, H - 1)
    iw = torch.clamp((cx * W).long(), 0, W - 1)
    picked = outputs[torch.arange(B, device=outputs.device), :, ih, iw].float()
    valid = (labels >= 0).all(dim=1)
    clean = labels.clone()
    clean[~valid] = 0.0
    tgt = torch.cat([clean, valid.float().unsqueeze(1)], dim=1)
    cls = F.binary_cross_entropy_with_logits(picked[:, 4], tgt[:, 4])
    pos = tgt[:, 4] > 0.5
    box = F.smooth_l1_loss(picked[:, :4][pos], tgt[:, :4][pos]) if pos.any() else picked.new_tensor(0.0)
    return cls_w * cls + box_w * box


_DICE = smp_heads.DiceLoss(mode="multiclass")


def task_loss(task_name, outputs, labels):
    if task_name == "segmentation":
        return _DICE(outputs.float(), labels)
    if task_name == "classification":
        return F.cross_entropy(outputs.float(), labels)
    if task_name == "detection":
        return detection_loss(outputs, labels)
    return F.mse_loss(outputs.float(), labels)


def synthetic_batch(task_cfg, batch, image_size, gen, device="cpu"):
    """Synthetic ultrasound-shaped batch for one task."""
    name, n = task_cfg["task_name"], task_cfg["num_classes"]
    x = torch.randn(batch, 3, image_size, image_size, generator=gen)
    if name == "segmentation":
        y = torch.randint(0, n, (batch, image_size, image_size), generator=gen)
    elif name == "classification":
        y = torch.randint(0, n, (batch,), generator=gen)
    elif name == "detection":
        a, b = torch.rand(batch, 2, generator=gen), torch.rand(batch, 2, generator=gen)
        lo, hi = torch.minimum(a, b), torch.maximum(a, b) + 1e-3
        y = torch.stack([lo[:, 0], lo[:, 1], hi[:, 0].clamp(max=1.0), hi[:, 1].clamp(max=1.0)], dim=1)
    else:
        y = torch.rand(batch, 2 * n, generator=gen)
    return x.to(device), y.to(device)


def train_step(model, optimizer, x, y, task_id, clip=1.0):
    """zero_grad -> forward -> loss -> backward -> clip -> step (train.py:326,440-455)."""
    name = model.task_id_to_name[task_id]
    out = model(x, task_id)
    loss = task_loss(name, out, y)
    optimizer.zero_grad()
    loss.backward()
    if clip > 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
    optimizer.step()
    return loss.detach()
