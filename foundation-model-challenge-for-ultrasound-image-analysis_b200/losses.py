"""Loss functions needed to drive a realistic backward through the hot path (PyTorch; out of kernel scope,
SURVEY §2 row 11).  Semantics follow ``/root/reference/code/losses/loss_functions.py`` for the losses
``configs/swin_b.yaml`` selects: smp multiclass Dice (:169), CrossEntropy (:176), the grid ``DetectionLoss``
(:10-53, with the centre-cell gather of code/train.py:395-418) and MSE (:199).
"""

import torch
import torch.nn as nn
import torch.nn.functional as F


class DiceLoss(nn.Module):
    """smp DiceLoss(mode='multiclass'): softmax probabilities, one-hot targets, sums over (batch, pixels),
    smooth 0, eps 1e-7, classes absent from the target masked out, mean over classes."""

    def __init__(self, mode="multiclass", eps=1e-7):
        super().__init__()
        if mode != "multiclass":
            raise NotImplementedError("only mode='multiclass' is used by the reference configs")
        self.eps = eps

    def forward(self, logits, target):
        bs, nc = logits.shape[0], logits.shape[1]
        prob = logits.float().log_softmax(dim=1).exp().view(bs, nc, -1)
        onehot = F.one_hot(target.view(bs, -1).long(), nc).permute(0, 2, 1).to(prob.dtype)
        inter = (prob * onehot).sum(dim=(0, 2))
        card = (prob + onehot).sum(dim=(0, 2))
        loss = 1.0 - (2.0 * inter) / card.clamp_min(self.eps)
        return (loss * (onehot.sum(dim=(0, 2)) > 0).to(loss.dtype)).mean()


class DetectionLoss(nn.Module):
    def __init__(self, classification_weight=2.0, box_regression_weight=1.0):
        super().__init__()
        self.cls_w, self.box_w = classification_weight, box_regression_weight

    def forward(self, predictions, targets):
        obj = F.binary_cross_entropy_with_logits(predictions[:, 4], targets[:, 4])
        pos = targets[:, 4] > 0.5
        if pos.any():
            box = F.smooth_l1_loss(predictions[:, :4][pos], targets[:, :4][pos])
        else:
            box = predictions.new_tensor(0.0)
        return self.cls_w * obj + self.box_w * box


def detection_targets(outputs, labels):
    """code/train.py:395-418 without the per-sample Python loop / host syncs: gather the prediction at the
    grid cell under the ground-truth box centre and build [bbox(4), objectness(1)] targets."""
    B, C, H, W = outputs.shape
    cx = (labels[:, 0] + labels[:, 2]) / 2.0
    cy = (labels[:, 1] + labels[:, 3]) / 2.0
    ih = torch.clamp((cy * H).long(), 0, H - 1)
    iw = torch.clamp((cx * W).long(), 0, W - 1)
    picked = outputs[torch.arange(B, device=outputs.device), :, ih, iw].float()
    valid = (labels >= 0).all(dim=1)
    clean = torch.where(valid[:, None], labels, torch.zeros_like(labels))
    return picked, torch.cat([clean, valid.float()[:, None]], dim=1)


def build_all_losses(config):
    fns = {}
    names = {t["task_name"] for t in config.get_task_configs()}
    for name in names:
        lc = config.get_loss_config(name)
        if name == "segmentation":
            fns[name] = nn.CrossEntropyLoss() if lc.get("type") == "CrossEntropyLoss" else DiceLoss("multiclass")
        elif name == "classification":
            fns[name] = nn.CrossEntropyLoss()
        elif name == "detection":
            if lc.get("type", "CenterNet").lower() == "centernet":
                raise NotImplementedError("CenterNet loss needs the CenterNet head (outside the shipped head family); "
                                          "set training.loss_configs.detection.type: Detection (SURVEY §8d caveat)")
            fns[name] = DetectionLoss(float(lc.get("classification_weight", 2.0)), float(lc.get("box_regression_weight", 1.0)))
        elif name == "Regression":
            fns[name] = {"L1Loss": nn.L1Loss, "SmoothL1Loss": nn.SmoothL1Loss}.get(lc.get("type"), nn.MSELoss)()
        else:
            raise ValueError(f"Unknown task name: {name}")
    weights = {k: float(v) for k, v in (config.get("training.loss_weights", {}) or {}).items()}
    return fns, weights


def compute_task_loss(loss_functions, task_name, outputs, labels):
    if task_name == "detection":
        picked, targets = detection_targets(outputs, labels)
        return loss_functions[task_name](picked, targets)
    if task_name in ("classification", "Regression"):
        return loss_functions[task_name](outputs.float(), labels)
    return loss_functions[task_name](outputs, labels)
