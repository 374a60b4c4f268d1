"""Configuration object with the reference's interface (``/root/reference/code/configs/__init__.py:9-117``).

Same dotted ``get(key, default)``, ``get_task_configs()``, ``set_task_configs_from_dataset()`` and
``get_loss_config()``; additionally constructible from a dict, and ``swin_b_27task()`` returns the
headline benchmark configuration (``configs/swin_b.yaml`` with the two changes SURVEY §8d requires:
``model.encoder.pretrained: null`` and ``training.loss_configs.detection.type: Detection``).
"""

import copy
from typing import Any, Dict, List, Optional

_SEG = ["T2A_fetal_abdomen", "T2A_fetal_brain", "T2A_fetal_femur", "T2A_fetal_thorax",
        "T2B_adult_liver_segment_2", "T2B_adult_liver_segment_3", "T2B_adult_liver_segment_4a",
        "T2B_adult_liver_segment_5", "T2B_adult_liver_segment_6", "T2B_adult_liver_segment_7",
        "T2B_adult_liver_segment_8", "T2C_fetal_head"]
_CLS = [("T1_fetal_planes", 6), ("T3A_breast_lymph_nodes", 2), ("T3A_breast_tumor", 2), ("T3B_liver_injury", 2),
        ("T3B_liver_steatosis", 2), ("T3C_thyroid_nodule", 2), ("T3D_liver_cirrhosis", 2), ("T3D_liver_fibrosis", 2),
        ("T3E_thyroid_cancer", 2)]
_DET = ["T4A_fetal_abdomen", "T4A_fetal_brain", "T4A_fetal_femur"]
_REG = ["T5_fetal_abdomen", "T5_fetal_brain", "T5_fetal_femur"]


def tasks_27() -> List[Dict]:
    """The 27 tasks of configs/swin_b.yaml:159-265 (12 seg, 9 cls, 3 det, 3 reg)."""
    t = [{"task_id": n, "task_name": "segmentation", "num_classes": 2} for n in _SEG]
    t += [{"task_id": n, "task_name": "classification", "num_classes": k} for n, k in _CLS]
    t += [{"task_id": n, "task_name": "detection", "num_classes": 1} for n in _DET]
    t += [{"task_id": n, "task_name": "Regression", "num_classes": 4} for n in _REG]
    return t


class Config:
    """Drop-in for the reference ``Config``: YAML path or dict in, dotted ``get`` out."""

    def __init__(self, source: Any = None):
        if isinstance(source, dict):
            self.config = copy.deepcopy(source)
        else:
            import yaml
            with open(source, "r", encoding="utf-8") as f:
                self.config = yaml.safe_load(f)

    def get(self, key: str, default: Any = None) -> Any:
        value = self.config
        try:
            for k in key.split("."):
                value = value[k]
            return value
        except (KeyError, TypeError):
            return default

    def get_task_configs(self) -> List[Dict]:
        return self.config["tasks"]

    def set_task_configs_from_dataset(self, task_configs: List[Dict]):
        self.config["tasks"] = task_configs
        self.config.setdefault("runtime", {})["tasks_from_dataset"] = True

    def tasks_from_dataset(self) -> bool:
        return bool(self.get("runtime.tasks_from_dataset", False))

    def get_loss_config(self, task_name: str) -> Dict:
        return self.config["training"]["loss_configs"].get(task_name, {})

    def __repr__(self):
        return f"Config(encoder={self.get('model.encoder.name')})"


def make_config(encoder: str = "swin_b", image_size: int = 224, batch_size: int = 32, tasks: Optional[List[Dict]] = None,
                separate_fpn: bool = True, dropout: float = 0.2, merge_policy: str = "cat",
                mixed_precision: bool = True) -> Config:
    """configs/swin_b.yaml as a dict (hot-path and head keys only) with the stated overrides."""
    return Config({
        "experiment": {"name": "mtus_b200", "seed": 42, "output_dir": "outputs"},
        "data": {"batch_size": batch_size, "image_size": image_size},
        "model": {
            "moe": {"enabled": False},
            "encoder": {"name": encoder, "pretrained": None, "freeze_encoder": False},
            "decoder": {"type": "fpn", "pyramid_channels": 256, "segmentation_channels": 128, "dropout": dropout,
                        "merge_policy": merge_policy, "separate_detection_fpn": separate_fpn,
                        "separate_classification_fpn": separate_fpn, "separate_regression_fpn": separate_fpn,
                        "use_fpn_for_classification": False, "use_fpn_for_regression": False},
            "use_film": False,
            "heads": {"segmentation": {"type": "baseline", "upsampling": 4, "mid_channels": 128, "num_blocks": 2,
                                       "use_deep_supervision": False},
                      "classification": {"type": "baseline", "mid_channels": 256, "dropout": 0.3},
                      "detection": {"type": "baseline", "mid_channels": 128, "num_anchors": 1},
                      "regression": {"type": "baseline", "hidden_dims": [256, 128], "use_tanh": True,
                                     "mid_channels": 256, "dropout": 0.3}},
        },
        "training": {
            "optimizer": {"type": "AdamW", "learning_rate": 1e-4, "weight_decay": 1e-4, "use_grouped_lr": True,
                          "encoder_lr_multiplier": 0.1, "head_lr_multiplier": 1.0},
            "loss_weights": {"segmentation": 1.0, "classification": 1.0, "detection": 1.0, "regression": 1.0},
            "loss_configs": {"segmentation": {"type": "DiceLoss", "mode": "multiclass"},
                             "classification": {"type": "CrossEntropyLoss"},
                             "detection": {"type": "Detection", "classification_weight": 2.0, "box_regression_weight": 1.0},
                             "regression": {"type": "MSELoss"}},
            "gradient_clip": 1.0,
        },
        "device": {"use_cuda": True, "multi_gpu": False, "mixed_precision": mixed_precision},
        "tasks": tasks if tasks is not None else tasks_27(),
    })


def swin_b_27task(batch_size: int = 32, image_size: int = 224, mixed_precision: bool = True) -> Config:
    return make_config("swin_b", image_size, batch_size, mixed_precision=mixed_precision)
