"""Input-level task prompt ("optional TaskPrompt2D add").

Restates ``/root/reference/code/models/task_prompt.py``: ``build_task_prompt_metadata`` (:28-72, a multi-hot descriptor per task
made of its task type, its ``num_classes_<n>`` tag and the tokens of its id) and ``TaskPrompt2D`` (:75-143: descriptor -> Linear ->
``[channels, p, p]`` map -> tanh -> bilinear resize (``align_corners=False``) -> ``x + s * prompt`` or ``x * (1 + s * prompt)``),
with the reference's parameter / buffer names (``task_prompt.prompt_proj.*``, ``task_prompt.prompt_scale``,
``task_prompt.task_metadata``) so checkpoints interchange.

The module sits in FRONT of the encoder (``multitask_model.py:198-199``), so training it needs d(loss)/d(image) out of the encoder:
``mtus_swin_input_grad`` (csrc/swin_exec.cu) provides it and ``_SwinFn.backward`` returns it to autograd.  The prompt itself is a
handful of tiny PyTorch ops (one ``[1, dim] x [dim, channels * p * p]`` product per step and one elementwise pass over the image
batch); the descriptor row is projected ONCE per call and broadcast over the batch instead of being expanded to ``B`` identical rows.
"""

import re
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

_PREFIX = re.compile(r"t\d+[a-z]?", re.IGNORECASE)      # "T2A", "t11": the running number of a task id carries no meaning


def _id_tokens(task_id) -> List[str]:
    toks = (t.strip().lower() for t in str(task_id).split("_"))
    return [t for t in toks if t and not _PREFIX.fullmatch(t)]


def build_task_prompt_metadata(task_configs: Sequence[Dict]) -> Tuple[torch.Tensor, Dict[str, int], Dict[str, List[str]]]:
    """-> (``[num_tasks, dim]`` multi-hot table, task_id -> row, vocabularies).  Column blocks, each sorted: task types,
    ``num_classes_<n>`` tags, id tokens."""
    ids = [str(c["task_id"]) for c in task_configs]
    types = [str(c.get("task_name", "unknown")).lower() for c in task_configs]
    tags = ["num_classes_%d" % int(c.get("num_classes", -1)) for c in task_configs]
    tokens = [_id_tokens(i) for i in ids]
    vocab_types, vocab_tags = sorted(set(types)), sorted(set(tags))
    vocab_tokens = sorted({t for ts in tokens for t in ts})
    off_tag, off_tok = len(vocab_types), len(vocab_types) + len(vocab_tags)
    table = torch.zeros(len(ids), off_tok + len(vocab_tokens), dtype=torch.float32)
    for row, (ty, tag, ts) in enumerate(zip(types, tags, tokens)):
        table[row, vocab_types.index(ty)] = 1.0
        table[row, off_tag + vocab_tags.index(tag)] = 1.0
        for t in ts:
            table[row, off_tok + vocab_tokens.index(t)] = 1.0
    info = {"task_types": vocab_types, "num_classes_tags": vocab_tags, "task_tokens": vocab_tokens}
    return table, {t: i for i, t in enumerate(ids)}, info


class TaskPrompt2D(nn.Module):
    def __init__(self, task_configs: Sequence[Dict], out_channels: int = 1, prompt_size: int = 32, inject_mode: str = "add",
                 init_scale: float = 0.1, use_tanh: bool = True):
        super().__init__()
        if inject_mode not in ("add", "mul"):
            raise ValueError(f"Unsupported inject_mode: {inject_mode}")
        table, self.task_id_to_idx, self.vocab_info = build_task_prompt_metadata(task_configs)
        if table.numel() == 0:
            raise ValueError("TaskPrompt2D received empty task metadata.")
        self.out_channels, self.prompt_size = int(out_channels), int(prompt_size)
        self.inject_mode, self.use_tanh = inject_mode, bool(use_tanh)
        self.register_buffer("task_metadata", table, persistent=True)
        self.prompt_proj = nn.Linear(table.shape[1], self.out_channels * self.prompt_size ** 2)
        self.prompt_scale = nn.Parameter(torch.tensor(float(init_scale), dtype=torch.float32))

    @property
    def prompt_dim(self) -> int:
        return int(self.task_metadata.shape[1])

    def _row(self, task_id, device) -> torch.Tensor:
        if task_id not in self.task_id_to_idx:
            raise ValueError(f"Unknown task_id for TaskPrompt2D: {task_id}")
        return self.task_metadata[self.task_id_to_idx[task_id]].to(device=device)

    def _map(self, task_id, spatial_size, device) -> torch.Tensor:
        """The prompt of one task as a ``[1, channels, H, W]`` map (identical for every image of the batch)."""
        p = self.prompt_proj(self._row(task_id, device).unsqueeze(0)).view(1, self.out_channels, self.prompt_size, self.prompt_size)
        if self.use_tanh:
            p = torch.tanh(p)
        if tuple(p.shape[-2:]) != tuple(spatial_size):
            p = F.interpolate(p, size=tuple(spatial_size), mode="bilinear", align_corners=False)
        return p

    def forward(self, task_id: str, batch_size: int, spatial_size: Tuple[int, int], device, return_vec: bool = False):
        prompt = self._map(task_id, spatial_size, device).expand(batch_size, -1, -1, -1)
        if return_vec:
            return prompt, self._row(task_id, device).unsqueeze(0).expand(batch_size, -1)
        return prompt

    def apply(self, x, task_id=None):
        """``apply(x, task_id)`` injects the prompt into the image batch (the reference's method name, task_prompt.py:132);
        ``apply(fn)`` keeps ``nn.Module.apply`` working (weight-init hooks walk every submodule with it)."""
        if task_id is None and callable(x) and not torch.is_tensor(x):
            return super().apply(x)
        p = self._map(task_id, x.shape[-2:], x.device).to(dtype=x.dtype)
        s = self.prompt_scale.to(dtype=x.dtype)
        return x + s * p if self.inject_mode == "add" else x * (1.0 + s * p)
