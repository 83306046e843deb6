"""Abstract vector quantizer -- same surface as the reference's BaseVectorQuantizer
(vqvae/modules/abstract_modules/base_quantizer.py:6-102); gather / argmin run in libvqgan_b200 kernels."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
from torch import nn

from ... import ops


class BaseVectorQuantizer(ABC, nn.Module):

    def __init__(self, num_embeddings: int, embedding_dim: int):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.codebook = nn.Embedding(self.num_embeddings, self.embedding_dim)     # base_quantizer.py:21
        self.kl_warmup = None
        self.temp_decay = None
        # data-parallel hook: callable(tensor) that sum-all-reduces EMA cluster statistics in place, plus the
        # number of ranks (so that Laplace smoothing sees the GLOBAL image batch)
        self.stats_allreduce = None
        self.world_size = 1
        self._prep = ops.CodebookPrep()           # cached bf16 split of the codebook for the fused search kernel

    def init_codebook(self) -> None:
        """uniform U(-1/K, 1/K) (base_quantizer.py:27-31)"""
        nn.init.uniform_(self.codebook.weight, -1 / self.num_embeddings, 1 / self.num_embeddings)

    @abstractmethod
    def forward(self, x: torch.Tensor):
        """x [B,D,H,W] -> (quantized [B,D,H,W], detached codes [B,H*W] int64, latent loss)"""

    @abstractmethod
    def vec_to_codes(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,D,H,W] -> flat codebook indices [B,H*W]"""

    @torch.no_grad()
    def get_codebook(self) -> torch.Tensor:
        return self.codebook.weight

    @torch.no_grad()
    def codes_to_vec(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [B,N] int -> [B,N,D] (base_quantizer.py:53-61)"""
        return ops.vq_gather(self.get_codebook(), codes)

    def get_codebook_usage(self, index_count: torch.Tensor):
        """index_count [K] -> (probabilities, perplexity, % used codes)   (base_quantizer.py:63-79).
        Epoch-level bookkeeping on K scalars: plain tensor arithmetic, not part of the per-step kernel path."""
        normalized = index_count / torch.sum(index_count)
        perplexity = torch.exp(-torch.sum(normalized * torch.log(normalized + 1e-10), dim=-1)).sum().item()
        used = torch.count_nonzero(normalized).item() * 100 / index_count.shape[0]
        return normalized, perplexity, used

    @torch.no_grad()
    def reduce_usage(self, index_count: torch.Tensor) -> torch.Tensor:
        """Sum the per-rank code-usage counts over the data-parallel group (no-op for one process).  The reference re-initialises
        from each rank's own last batch, so its replicas diverge; here every rank sees the same global counts."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            index_count = index_count.clone()
            dist.all_reduce(index_count, op=dist.ReduceOp.SUM)
        return index_count

    @torch.no_grad()
    def reinit_unused_codes(self, codebook_usage: torch.Tensor):
        """Re-sample dead codes from live ones (base_quantizer.py:81-102); epoch-level, K-sized.  Data parallel: the multinomial
        draw happens on rank 0 only and is broadcast, so that all replicas keep identical codebooks (`codebook_usage` must be the
        same on every rank: see reduce_usage)."""
        import torch.distributed as dist
        unused = torch.nonzero(codebook_usage == 0.).squeeze(1)
        n_unused = unused.shape[0]
        if n_unused > 0:
            multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            if not multi or dist.get_rank() == 0:
                det = torch.are_deterministic_algorithms_enabled()
                torch.use_deterministic_algorithms(False)
                sampled = torch.multinomial(codebook_usage, n_unused, replacement=True)
                torch.use_deterministic_algorithms(det)
            else:
                sampled = torch.empty(n_unused, dtype=torch.int64, device=codebook_usage.device)
            if multi:
                dist.broadcast(sampled, src=0)
            self.codebook.weight[unused] = self.codebook.weight[sampled].clone()
            ops.bump_weights_epoch()
