"""Per-task FiLM parameters for the decoder output.

Restates the two generators of ``/root/reference/code/models/film_layer.py`` (``TaskFiLMGenerator`` :103-148, one learnable
(gamma, beta) pair per task; ``TaskEmbeddingFiLMGenerator`` :151-216, task embedding -> two small MLPs) with the reference's
parameter names, so ``film_generator.*`` checkpoint keys load unchanged.  The modulation itself
(``FiLMLayer.forward``, :94-99: ``gamma * x + beta`` over the ``[B, C, H/4, W/4]`` decoder output, called at
``multitask_model.py:214-216, 224-226``) is NOT a PyTorch pass here: ``FPNDecoder.forward(features, film=(gamma, beta))``
applies it inside the FPN merge kernel (``mtus_fpn_merge_film_fwd``) and ``mtus_film_grad`` produces dgamma / dbeta.
"""

import torch
import torch.nn as nn


class TaskFiLMGenerator(nn.Module):
    def __init__(self, task_ids, num_features, use_affine=True):
        super().__init__()
        self.task_ids, self.num_features, self.use_affine = list(task_ids), int(num_features), bool(use_affine)
        self.task_gammas = nn.ParameterDict({t: nn.Parameter(torch.ones(self.num_features)) for t in self.task_ids})
        if self.use_affine:
            self.task_betas = nn.ParameterDict({t: nn.Parameter(torch.zeros(self.num_features)) for t in self.task_ids})

    def forward(self, task_id):
        if task_id not in self.task_gammas:
            raise ValueError(f"Unknown task_id: {task_id}. Available: {self.task_ids}")
        return self.task_gammas[task_id], (self.task_betas[task_id] if self.use_affine else None)


def _mlp(d_in, d_out):
    return nn.Sequential(nn.Linear(d_in, 2 * d_out), nn.ReLU(), nn.Linear(2 * d_out, d_out))


class TaskEmbeddingFiLMGenerator(nn.Module):
    def __init__(self, task_ids, num_features, embedding_dim=64, use_affine=True):
        super().__init__()
        self.task_ids, self.num_features, self.use_affine = list(task_ids), int(num_features), bool(use_affine)
        self.task_id_to_idx = {t: i for i, t in enumerate(self.task_ids)}
        self.task_embeddings = nn.Embedding(len(self.task_ids), embedding_dim)
        self.gamma_generator = _mlp(embedding_dim, self.num_features)
        if self.use_affine:
            self.beta_generator = _mlp(embedding_dim, self.num_features)

    def forward(self, task_id):
        if task_id not in self.task_id_to_idx:
            raise ValueError(f"Unknown task_id: {task_id}")
        emb = self.task_embeddings.weight[self.task_id_to_idx[task_id]]
        return self.gamma_generator(emb), (self.beta_generator(emb) if self.use_affine else None)


class FiLMLayer(nn.Module):
    """Parameter-free placeholder keeping the reference's module tree (``film_layer``); the arithmetic is fused into the
    decoder's merge kernel.  Calling it on a tensor (a caller outside MultiTaskModel) applies the same formula in PyTorch."""

    def __init__(self, num_features, use_affine=True):
        super().__init__()
        self.num_features, self.use_affine = int(num_features), bool(use_affine)

    def forward(self, x, condition=None):
        gamma, beta = condition
        out = gamma.view(1, -1, 1, 1) * x
        return out + beta.view(1, -1, 1, 1) if (self.use_affine and beta is not None) else out


def build_film(config, task_ids, num_features):
    """multitask_model.py:52-79: ``model.film.{use_task_embedding, embedding_dim, use_affine}``."""
    fc = config.get("model.film", {}) or {}
    use_affine = bool(fc.get("use_affine", True))
    if fc.get("use_task_embedding", False):
        gen = TaskEmbeddingFiLMGenerator(task_ids, num_features, int(fc.get("embedding_dim", 64)), use_affine)
    else:
        gen = TaskFiLMGenerator(task_ids, num_features, use_affine)
    return gen, FiLMLayer(num_features, use_affine)
