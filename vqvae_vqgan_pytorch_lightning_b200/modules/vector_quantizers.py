"""The four vector quantizers of the reference (vqvae/modules/vector_quantizers.py), same constructors and
return contracts, computed by the fused libvqgan_b200 VQ kernels (no N x K distance / one-hot matrices in HBM)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from .abstract_modules.base_quantizer import BaseVectorQuantizer


class VectorQuantizer(BaseVectorQuantizer):
    """Standard VQ-VAE quantizer (vector_quantizers.py:8-84): loss = mse(q, z.detach()) + beta * mse(q.detach(), z)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float = 0.25):
        super().__init__(num_embeddings, embedding_dim)
        self.commitment_cost = commitment_cost

    def forward(self, x: torch.Tensor):
        q, idx, loss, _, _ = ops.vq_quantize(x, self.codebook.weight, 0, self.commitment_cost, 1.0, False)
        return q, idx, loss

    @torch.no_grad()
    def vec_to_codes(self, x: torch.Tensor) -> torch.Tensor:
        return ops.vq_codes(x, self.codebook.weight, 0)


class EMAVectorQuantizer(BaseVectorQuantizer):
    """EMA codebook (vector_quantizers.py:87-203).  The codebook is frozen for the optimizer (:114) and replaced by
    ema_weight / ema_count after every training forward; Laplace smoothing uses the IMAGE batch size (defect B7,
    replicated).  Under data parallelism the cluster statistics (counts, sums) are all-reduced first and the
    smoothing sees the global batch, which equals the reference run single-process on the concatenated batch."""

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float = 0.25, decay: float = 0.95,
                 epsilon: float = 1e-5):
        super().__init__(num_embeddings, embedding_dim)
        self.commitment_cost = commitment_cost
        self.codebook.requires_grad_(False)
        self.codebook.weight.requires_grad_(False)
        self.register_buffer('ema_count', torch.zeros(self.num_embeddings))
        self.register_buffer('ema_weight', torch.empty((self.num_embeddings, self.embedding_dim)))
        self.ema_weight.uniform_(-1 / self.num_embeddings, 1 / self.num_embeddings)
        self.decay = decay
        self.epsilon = epsilon

    def forward(self, x: torch.Tensor):
        q, idx, loss, counts, dw = ops.vq_quantize(x, self.codebook.weight, 0, self.commitment_cost, 0.0, self.training)
        if self.training:
            with torch.no_grad():
                batch = x.shape[0]
                if self.stats_allreduce is not None:
                    self.stats_allreduce(counts, dw)
                    batch = batch * self.world_size
                ops.vq_ema_update(self.ema_count, self.ema_weight.data, self.codebook.weight.data, counts, dw, self.decay,
                                  self.epsilon, batch)
                ops.bump_weights_epoch()
        return q, idx, loss

    @torch.no_grad()
    def vec_to_codes(self, x: torch.Tensor) -> torch.Tensor:
        return ops.vq_codes(x, self.codebook.weight, 0)
