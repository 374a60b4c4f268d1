"""B200-native (sm_100a) implementation of MTUS-Net's shared-encoder hot path: Swin Transformer encoder +
FPN decoder, forward and backward, behind the reference's own Python API.

Import name: ``mtus_b200`` (see ``mtus_b200.py`` at the repository root -- the directory name mandated for
this package contains hyphens and cannot be imported directly).

Public surface (mirrors /root/reference/code/models/__init__.py:3-16 for the hot path):
    build_encoder, SwinTransformerEncoder, SWIN_MODEL_MAPPING      (encoders.py)
    build_fpn_decoder, build_decoders, FPNDecoder                  (decoders.py)
    MultiTaskModel, build_model, build_optimizer                   (multitask_model.py)
    Config, swin_b_27task, make_config                             (configs.py)
    DistributedTaskSampler, GradAllReducer, DataParallelTrainer, DevicePrefetcher    (parallel.py)
"""

from . import _lib
from .configs import Config, make_config, swin_b_27task, tasks_27
from .encoders import build_encoder, SwinTransformerEncoder, SwinCore, SWIN_MODEL_MAPPING, SWIN_ARCHS
from .decoders import build_fpn_decoder, build_decoders, FPNDecoder
from .heads import build_all_heads, build_task_head
from .losses import build_all_losses, compute_task_loss
from .multitask_model import MultiTaskModel, build_model, build_optimizer, build_flat_optimizer
from .optim import FlatAdamW
from .checkpoint import load_pretrained, load_checkpoint, convert_swin_state_dict, resize_rel_pos_bias_table
from .parallel import DistributedTaskSampler, DistributedEvalBatchSampler, gather_task_metrics, GradAllReducer, DataParallelTrainer, DevicePrefetcher, synthetic_batch

__all__ = [
    "build_encoder", "SwinTransformerEncoder", "SwinCore", "SWIN_MODEL_MAPPING", "SWIN_ARCHS",
    "build_fpn_decoder", "build_decoders", "FPNDecoder", "build_all_heads", "build_task_head",
    "build_all_losses", "compute_task_loss", "MultiTaskModel", "build_model", "build_optimizer", "build_flat_optimizer", "FlatAdamW",
    "load_pretrained", "load_checkpoint", "convert_swin_state_dict", "resize_rel_pos_bias_table",
    "Config", "make_config", "swin_b_27task", "tasks_27",
    "DistributedTaskSampler", "DistributedEvalBatchSampler", "gather_task_metrics", "GradAllReducer", "DataParallelTrainer", "DevicePrefetcher", "synthetic_batch",
]
