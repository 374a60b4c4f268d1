"""``MultiTaskModel`` with the reference's constructor, ``forward(x, task_id)`` signature and routing
(``/root/reference/code/models/multitask_model.py:13-362``), built on the native encoder / FPN.

Kept: ``build_model(config)`` (:346), ``forward`` routing by task type incl. the ``ValueError`` on an
unknown id (:187-188), ``use_fpn_for_{cls,reg}`` (:45-46), decoder aliasing (:294-303),
``get_trainable_parameters`` (:282), ``freeze_encoder/unfreeze_encoder`` (:333-343),
``get_moe_aux_loss/get_moe_stats`` (:310-331), attribute names (``encoder``, ``fpn_decoder_{seg,det,cls,reg}``,
``heads``) and therefore the checkpoint keys.  FiLM on the decoder output (``model.use_film``, :52-79, 214-216) is provided with the
modulation fused into the decoder's merge kernel; the input-level TaskPrompt2D (``model.task_prompt``, :81-111,
194-199) is a small PyTorch module in front of the encoder whose gradient the encoder returns (``mtus_swin_input_grad``, §8f N4).
Not provided (outside the hot path): MoE -- requesting it raises.
"""

import torch
import torch.nn as nn

from .encoders import build_encoder
from .decoders import build_decoders
from .heads import build_all_heads


class MultiTaskModel(nn.Module):
    def __init__(self, config, precision=None, zero_copy_features=True):
        super().__init__()
        self.config = config
        self.task_configs = config.get_task_configs()
        if config.get("model.moe.enabled", False):
            raise NotImplementedError("mtus_b200: model.moe.enabled is outside the hot path and not provided")
        task_ids = [c["task_id"] for c in self.task_configs]
        self.encoder = build_encoder(config, task_ids=task_ids, precision=precision, zero_copy_features=zero_copy_features)
        self.precision = self.encoder.model.precision
        encoder_channels = list(self.encoder.out_channels)
        decoders = build_decoders(self.encoder, config, precision=self.precision)
        self.fpn_decoder_seg = decoders["fpn_seg"]
        self.fpn_decoder_det = decoders["fpn_det"]
        self.fpn_decoder_cls = decoders["fpn_cls"]
        self.fpn_decoder_reg = decoders["fpn_reg"]
        self.use_fpn_for_cls = config.get("model.decoder.use_fpn_for_classification", True)
        self.use_fpn_for_reg = config.get("model.decoder.use_fpn_for_regression", True)
        self.fpn_out_channels = self.fpn_decoder_seg.out_channels
        # FiLM on the decoder output (multitask_model.py:52-79, 214-216): generators as in the reference, the modulation
        # itself fused into the decoder's merge kernel
        self.use_film = bool(config.get("model.use_film", False))
        if self.use_film:
            from .film import build_film
            self.film_generator, self.film_layer = build_film(config, task_ids, self.fpn_out_channels)
        # input-level task prompt (multitask_model.py:81-111): a small PyTorch module in front of the encoder; its gradient
        # comes out of the encoder through mtus_swin_input_grad
        tp_cfg = config.get("model.task_prompt", {}) or {}
        self.use_task_prompt = bool(tp_cfg.get("enabled", False))
        names = tp_cfg.get("apply_to_task_names", None)
        self.task_prompt_apply_task_names = None if names is None else {str(n).lower() for n in names}
        if self.use_task_prompt:
            if hasattr(config, "tasks_from_dataset") and not config.tasks_from_dataset():
                raise ValueError("TaskPrompt2D requires dataset-derived task configs. "
                                 "Load dataset metadata and override config tasks before building the model.")
            from .task_prompt import TaskPrompt2D
            self.task_prompt = TaskPrompt2D(self.task_configs, out_channels=int(tp_cfg.get("channels", 1)),
                                            prompt_size=int(tp_cfg.get("prompt_size", 32)),
                                            inject_mode=str(tp_cfg.get("inject_mode", "add")).lower(),
                                            init_scale=float(tp_cfg.get("init_scale", 0.1)),
                                            use_tanh=bool(tp_cfg.get("use_tanh", True)))
        self.use_moe = False
        self.heads = build_all_heads(self.task_configs, self.fpn_out_channels, encoder_channels,
                                     config.config.get("model", {}) if hasattr(config, "config") else {})
        self.task_id_to_name = {c["task_id"]: c["task_name"] for c in self.task_configs}
        # the FPN hands channels-last tensors to the heads: keep their conv weights in the matching memory format
        self.heads.to(memory_format=torch.channels_last)

    def _head(self, task_id, x):
        # the heads are plain PyTorch modules with fp32 parameters: run them under autocast in bf16 mode
        if self.precision == "bf16":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.heads[task_id](x)
        return self.heads[task_id](x)

    def forward(self, x, task_id):
        if task_id not in self.heads:
            raise ValueError(f"Unknown task_id: {task_id}")
        task_name = self.task_id_to_name[task_id]
        if self.use_task_prompt and (self.task_prompt_apply_task_names is None
                                     or task_name.lower() in self.task_prompt_apply_task_names):
            if x.dtype == torch.uint8:
                raise TypeError("mtus_b200: TaskPrompt2D modulates the NORMALISED image; pass the float [B,3,H,W] batch, not "
                                "the raw uint8 one")
            x = self.task_prompt.apply(x, task_id)
        features = self.encoder(x, task_id) if getattr(self.encoder, "supports_task_id", False) else self.encoder(x)
        film = self.film_generator(task_id) if self.use_film else None
        if task_name == "segmentation":
            return self._head(task_id, self.fpn_decoder_seg(features, film=film))
        if task_name == "detection":
            return self._head(task_id, self.fpn_decoder_det(features, film=film))
        if task_name == "classification":
            if self.use_fpn_for_cls:
                return self._head(task_id, self.fpn_decoder_cls(features, film=film))
            return self._head(task_id, features)
        if self.use_fpn_for_reg:
            return self._head(task_id, self.fpn_decoder_reg(features, film=film))
        return self._head(task_id, features)

    def get_trainable_parameters(self):
        encoder_params = list(self.encoder.parameters())
        head_params = list(self.fpn_decoder_seg.parameters())
        if self.fpn_decoder_det is not self.fpn_decoder_seg:
            head_params += list(self.fpn_decoder_det.parameters())
        if self.use_fpn_for_cls and self.fpn_decoder_cls not in (self.fpn_decoder_seg, self.fpn_decoder_det):
            head_params += list(self.fpn_decoder_cls.parameters())
        if self.use_fpn_for_reg and self.fpn_decoder_reg not in (self.fpn_decoder_seg, self.fpn_decoder_det,
                                                                 self.fpn_decoder_cls):
            head_params += list(self.fpn_decoder_reg.parameters())
        head_params += list(self.heads.parameters())
        if self.use_task_prompt:
            head_params += list(self.task_prompt.parameters())
        # as in the reference (multitask_model.py:282-308) the FiLM generator's parameters are in NEITHER group: with
        # grouped learning rates (train.py:184-190) they keep their initial values; model.parameters() still lists them
        return encoder_params, head_params

    def get_moe_aux_loss(self):
        return torch.tensor(0.0, device=next(self.parameters()).device)

    def get_moe_stats(self):
        return []

    def freeze_encoder(self):
        for p in self.encoder.parameters():
            p.requires_grad = False

    def unfreeze_encoder(self):
        for p in self.encoder.parameters():
            p.requires_grad = True


def build_model(config, precision=None, zero_copy_features=True):
    model = MultiTaskModel(config, precision=precision, zero_copy_features=zero_copy_features)
    if config.get("model.encoder.freeze_encoder", False):
        model.freeze_encoder()
    return model


def build_flat_optimizer(model, config):
    """Same rule and grouping as ``build_optimizer`` + ``clip_grad_norm_(training.gradient_clip)``, but the encoder and
    the FPN decoders are updated with one kernel per flat parameter block (optim.FlatAdamW)."""
    from .optim import FlatAdamW
    return FlatAdamW(model, lr=float(config.get("training.optimizer.learning_rate", 1e-4)),
                     weight_decay=float(config.get("training.optimizer.weight_decay", 1e-4)),
                     encoder_lr_multiplier=float(config.get("training.optimizer.encoder_lr_multiplier", 0.1)),
                     head_lr_multiplier=float(config.get("training.optimizer.head_lr_multiplier", 1.0)),
                     max_grad_norm=float(config.get("training.gradient_clip", 1.0)))


def build_optimizer(model, config, fused=None):
    """AdamW with the reference's grouped learning rates (code/train.py:176-219): encoder x0.1, heads x1.0.
    ``fused=True`` selects torch's multi-tensor fused CUDA AdamW (same update rule)."""
    lr = float(config.get("training.optimizer.learning_rate", 1e-4))
    wd = float(config.get("training.optimizer.weight_decay", 1e-4))
    enc, head = model.get_trainable_parameters()
    groups = [{"params": enc, "lr": lr * float(config.get("training.optimizer.encoder_lr_multiplier", 0.1))},
              {"params": head, "lr": lr * float(config.get("training.optimizer.head_lr_multiplier", 1.0))}]
    return torch.optim.AdamW(groups, lr=lr, weight_decay=wd, fused=fused)
