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
        q, idx, loss, _ = ops.vq_quantize(x, self.codebook.weight, 0, self.commitment_cost, 1.0, False, self._prep)
        return q, idx, loss

    @torch.no_grad()
    def vec_to_codes(self, x: torch.Tensor) -> torch.Tensor:
        return ops.vq_codes(x, self.codebook.weight, 0, self._prep)


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
        q, idx, loss, stats = ops.vq_quantize(x, self.codebook.weight, 0, self.commitment_cost, 0.0, self.training, self._prep)
        if self.training:
            with torch.no_grad():
                batch = x.shape[0]
                if self.stats_allreduce is not None:
                    self.stats_allreduce(stats)             # ONE buffer [counts | dw]: a single all-reduce
                    batch = batch * self.world_size
                k = self.num_embeddings
                counts, dw = stats[:k], stats[k:].view(k, self.embedding_dim)
                ops.vq_ema_update(self.ema_count, self.ema_weight.data, self.codebook.weight.data, counts, dw, self.decay,
                                  self.epsilon, batch, prep=self._prep)
                self._prep.invalidate_unless_fresh(self.codebook.weight)       # (no global weights-epoch bump: only the codebook moved)
        return q, idx, loss

    @torch.no_grad()
    def vec_to_codes(self, x: torch.Tensor) -> torch.Tensor:
        return ops.vq_codes(x, self.codebook.weight, 0, self._prep)


class GumbelVectorQuantizer(BaseVectorQuantizer):
    """Gumbel-softmax quantizer (vector_quantizers.py:206-274).  The encoder emits K channels; logits = 1x1 conv K->K
    (implicit GEMM), y = gumbel_softmax(logits, tau) over channels (row kernel on the channels-last tensor), quantized =
    einsum('b n h w, n d -> b d h w', y, codebook) = a second 1x1 conv with the transposed codebook as weight, plus the
    KL-to-uniform term.  Defects replicated: noise is always added, even in eval mode (B6); indices are (B,H,W) (B5).
    `exp_noise` (the Exp(1) samples F.gumbel_softmax draws) may be passed explicitly for parity runs."""

    def __init__(self, num_embeddings: int, embedding_dim: int, straight_through: bool = False, temp: float = 1.0,
                 kl_cost: float = 5e-4):
        super().__init__(num_embeddings, embedding_dim)
        from .autoencoder import Conv2d
        self.x_to_logits = Conv2d(num_embeddings, num_embeddings, 1)
        self.straight_through = straight_through
        self.temp = temp
        self.kl_cost = kl_cost
        # device mirror of (temp, kl_cost) for CUDA-graph replay: uploaded by set_consts() (ops.StepScalars), read by the kernels
        self._consts = None
        self.device_consts = False

    def enable_device_consts(self, device) -> None:
        """(temp, kl_cost) reach the kernels through device memory from now on (the CUDA-graph trainer calls this BEFORE a step is
        captured: the captured kernels keep reading the buffer that set_consts() refreshes before every replay)."""
        if self._consts is None or self._consts.dev.device != torch.device(device):
            self._consts = ops.StepScalars((2,), device)
        self.device_consts = True
        self._upload_consts()

    def _upload_consts(self) -> None:
        self._consts.upload(torch.tensor([float(self.temp), float(self.kl_cost)], dtype=torch.float32))

    def forward(self, x: torch.Tensor, exp_noise: torch.Tensor = None):
        hard = self.straight_through if self.training else True
        logits = self.x_to_logits(x, out_dtype=torch.float32)
        if exp_noise is None:
            exp_noise = torch.empty_like(logits, memory_format=torch.preserve_format).exponential_()     # RNG plumbing
        if self.device_consts:
            # (temp, kl_cost) from device memory: the values set_consts() uploaded for this step (schedules of model.py:219-225)
            if self._consts is None or self._consts.dev.device != logits.device:
                raise RuntimeError('GumbelVectorQuantizer: enable_device_consts(device) was not called for this device')
            y, idx, kl_mean = ops.gumbel_rows(logits, exp_noise, self._consts.dev[:1], hard)
            kl = self._consts.dev[1] * kl_mean
        else:
            y, idx, kl_mean = ops.gumbel_rows(logits, exp_noise, float(self.temp), hard)
            kl = self.kl_cost * kl_mean
        w = self.codebook.weight.t().reshape(self.embedding_dim, self.num_embeddings, 1, 1)
        quantized = ops.conv2d(y, w, out_dtype=torch.float32)
        self.last_counts = None
        return quantized, idx, kl

    def get_consts(self):
        return self.temp, self.kl_cost

    def set_consts(self, temp: float = None, kl_cost: float = None) -> None:
        if temp is not None:
            self.temp = temp
        if kl_cost is not None:
            self.kl_cost = kl_cost
        if self._consts is not None:
            self._upload_consts()

    @torch.no_grad()
    def vec_to_codes(self, x: torch.Tensor, exp_noise: torch.Tensor = None) -> torch.Tensor:
        """gumbel_softmax(x, tau=1, hard=True).argmax(1) on the RAW encoder output (x_to_logits is skipped: B6)."""
        x = ops.as_nhwc(x, torch.float32)
        if exp_noise is None:
            exp_noise = torch.empty_like(x, memory_format=torch.preserve_format).exponential_()
        return ops.gumbel_rows(x, exp_noise, 1.0, True)[1]


class EntropyVectorQuantizer(BaseVectorQuantizer):
    """Entropy-regularised quantizer (vector_quantizers.py:277-381)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, ent_loss_ratio: float = 0.1, ent_temperature: float = 0.01,
                 ent_loss_type: str = 'softmax', commitment_cost: float = 0.25):
        super().__init__(num_embeddings, embedding_dim)
        self.ent_loss_ratio = ent_loss_ratio
        self.ent_temperature = ent_temperature
        self.ent_loss_type = ent_loss_type
        self.commitment_cost = commitment_cost

    def forward(self, x: torch.Tensor):
        if self.ent_loss_type not in ('softmax', 'argmax'):
            raise ValueError('Entropy loss {} not supported'.format(self.ent_loss_type))
        return ops.vq_entropy(x, self.codebook.weight, self.commitment_cost, self.ent_loss_ratio, self.ent_temperature,
                              argmax_target=(self.ent_loss_type == 'argmax'))

    @torch.no_grad()
    def vec_to_codes(self, x: torch.Tensor) -> torch.Tensor:
        return ops.vq_codes(x, self.codebook.weight, 1, self._prep)
