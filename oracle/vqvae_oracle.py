"""CPU oracle for the VQ-VAE / VQGAN training step  --  TEST INFRASTRUCTURE ONLY.

This file is a functional (state-dict driven) restatement, in plain fp32 PyTorch on the
CPU, of the arithmetic of the reference's hot path.  It is *not* part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs (cpu_baseline,
``--impl reference`` and ``torch_eager_gpu_baseline`` -- the same port timed as stock torch eager ops
with its tensors on the GPU; reported baselines, never the thing shipped) may import it.  The product path (``vqvae_vqgan_pytorch_lightning_b200``) never does
and fails loudly when its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, so this
oracle is pinned by *executing the reference's own modules* in the build container
(``oracle/make_golden.py`` imports ``/root/reference/vqvae/modules/*``) and committing the
outputs as fixtures in ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file
against those fixtures.  Third-party pieces with no pin (scheduling_utils, kornia) are
marked "parity unpinned" where they are restated.

Every function cites the reference file:line (relative to the reference repo root) that
it follows.  Tensors named by a flat dict ``sd`` use the reference's state_dict keys
(``encoder.blocks.0.norm1.weight`` ...), which is also how the product stores them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

GN_GROUPS = 32
GN_EPS = 1e-6


# --------------------------------------------------------------------------------------
# autoencoder pieces  (vqvae/modules/autoencoder.py)
# --------------------------------------------------------------------------------------
def group_norm(x: Tensor, weight: Tensor, bias: Tensor, groups: int = GN_GROUPS, eps: float = GN_EPS) -> Tensor:
    """autoencoder.py:25-39 -- per (sample, group) mean and UNBIASED variance (torch.var default),
    normalise, then per-channel affine with (1,C,1,1) parameters."""
    b, c, h, w = x.shape
    xg = x.reshape(b, groups, (c // groups) * h * w)
    mu = xg.mean(dim=2, keepdim=True)
    var = xg.var(dim=2, keepdim=True)  # unbiased: divides by n-1
    xn = ((xg - mu) / torch.sqrt(var + eps)).reshape(b, c, h, w)
    return xn * weight.reshape(1, c, 1, 1) + bias.reshape(1, c, 1, 1)


def res_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """autoencoder.py:63-77 -- GN,SiLU,3x3, GN,SiLU,3x3 (+1x1 shortcut on the raw input) + skip.
    All convs bias-free, 'same' padding."""
    h = F.silu(group_norm(x, sd[p + 'norm1.weight'], sd[p + 'norm1.bias']))
    h = F.conv2d(h, sd[p + 'conv1.weight'], None, padding=1)
    h = F.silu(group_norm(h, sd[p + 'norm2.weight'], sd[p + 'norm2.bias']))
    h = F.conv2d(h, sd[p + 'conv2.weight'], None, padding=1)
    if (p + 'conv_shortcut.weight') in sd:
        x = F.conv2d(x, sd[p + 'conv_shortcut.weight'], None)
    return x + h


def encoder_layout(num_res_blocks: int, channel_multipliers: Sequence[int]) -> List[Tuple[str, int]]:
    """Index layout of Encoder.blocks (autoencoder.py:116-128): per level num_res_blocks ResBlocks then a
    Downsample.  Returns [('res', idx) | ('down', idx)]."""
    out, i = [], 0
    for _ in channel_multipliers:
        for _ in range(num_res_blocks):
            out.append(('res', i)); i += 1
        out.append(('down', i)); i += 1
    return out


def encoder_forward(sd: SD, x: Tensor, num_res_blocks: int, channel_multipliers: Sequence[int],
                    prefix: str = 'encoder.') -> Tensor:
    """autoencoder.py:135-143."""
    x = F.conv2d(x, sd[prefix + 'conv_in.weight'], None, padding=1)
    for kind, i in encoder_layout(num_res_blocks, channel_multipliers):
        if kind == 'res':
            x = res_block(sd, f'{prefix}blocks.{i}.', x)
        else:
            x = F.avg_pool2d(x, 2, 2, 0)           # autoencoder.py:89-91
    for j in range(num_res_blocks):
        x = res_block(sd, f'{prefix}final_residual.{j}.', x)
    x = F.silu(group_norm(x, sd[prefix + 'norm.weight'], sd[prefix + 'norm.bias']))
    return F.conv2d(x, sd[prefix + 'conv_out.weight'], sd[prefix + 'conv_out.bias'])


def decoder_forward(sd: SD, x: Tensor, num_res_blocks: int, channel_multipliers: Sequence[int],
                    prefix: str = 'decoder.') -> Tensor:
    """autoencoder.py:172-180; Upsample = nearest-exact x2 then 3x3 conv with bias (:103-106)."""
    x = F.conv2d(x, sd[prefix + 'conv_in.weight'], sd[prefix + 'conv_in.bias'], padding=1)
    for j in range(num_res_blocks):
        x = res_block(sd, f'{prefix}initial_residual.{j}.', x)
    i = 0
    for _ in channel_multipliers:
        for _ in range(num_res_blocks):
            x = res_block(sd, f'{prefix}blocks.{i}.', x); i += 1
        x = F.interpolate(x, scale_factor=2.0, mode='nearest-exact')
        x = F.conv2d(x, sd[f'{prefix}blocks.{i}.conv.weight'], sd[f'{prefix}blocks.{i}.conv.bias'], padding=1)
        i += 1
    x = F.silu(group_norm(x, sd[prefix + 'norm.weight'], sd[prefix + 'norm.bias']))
    x = F.conv2d(x, sd[prefix + 'conv_out.weight'], sd[prefix + 'conv_out.bias'], padding=1)
    return torch.tanh(x)


# --------------------------------------------------------------------------------------
# vector quantizers  (vqvae/modules/vector_quantizers.py)
# --------------------------------------------------------------------------------------
def _flatten_latents(z: Tensor) -> Tensor:
    b, c, h, w = z.shape
    return z.permute(0, 2, 3, 1).reshape(b * h * w, c)   # 'b c h w -> (b h w) c'


def _unflatten(q: Tensor, shape) -> Tensor:
    b, c, h, w = shape
    return q.reshape(b, h, w, c).permute(0, 3, 1, 2)


def l2_distances(flat: Tensor, codebook: Tensor, order: str = 'standard') -> Tensor:
    """fp32 distance matrix with the reference's exact operation order.
    'standard'/'ema': (|z|^2 + |e|^2) - 2 z.e   (vector_quantizers.py:37-39, 142-144)
    'entropy'       : (|z|^2 - 2 z.e) + |e|^2   (vector_quantizers.py:337-340)"""
    a2 = torch.sum(flat ** 2, dim=1, keepdim=True)
    if order == 'entropy':
        cbt = codebook.T
        b2 = torch.sum(cbt ** 2, dim=0, keepdim=True)
        return a2 - 2 * torch.matmul(flat, cbt) + b2
    b2 = torch.sum(codebook ** 2, dim=1)
    return a2 + b2 - 2 * torch.matmul(flat, codebook.t())


def vq_standard(z: Tensor, codebook: Tensor, beta: float):
    """vector_quantizers.py:23-61.  Returns (quantized[B,D,H,W] with straight-through grad, idx[B,HW] i64, loss)."""
    flat = _flatten_latents(z)
    idx = torch.argmin(l2_distances(flat, codebook), dim=1)
    q = codebook[idx]                        # == one_hot @ codebook (exact row selection)
    e_loss = beta * F.mse_loss(q.detach(), flat)
    q_loss = F.mse_loss(q, flat.detach())
    q_st = flat + (q - flat).detach()
    return _unflatten(q_st, z.shape), idx.reshape(z.shape[0], -1).detach(), q_loss + e_loss


def vq_ema(z: Tensor, codebook: Tensor, ema_count: Tensor, ema_weight: Tensor, beta: float, decay: float,
           epsilon: float, training: bool = True, batch_for_smoothing: Optional[int] = None):
    """vector_quantizers.py:128-180.  Quantises with the OLD codebook, then (training) updates the EMA state
    with Laplace smoothing over b = number of IMAGES (defect B7, replicated).  Returns
    (quantized, idx, loss, new_codebook, new_ema_count, new_ema_weight).  ``batch_for_smoothing`` lets the
    data-parallel tests pass the GLOBAL batch."""
    b = z.shape[0] if batch_for_smoothing is None else batch_for_smoothing
    K = codebook.shape[0]
    flat = _flatten_latents(z)
    idx = torch.argmin(l2_distances(flat, codebook), dim=1)
    q = codebook[idx]
    new_cb, new_cnt, new_w = codebook, ema_count, ema_weight
    if training:
        with torch.no_grad():
            counts = torch.bincount(idx, minlength=K).to(ema_count.dtype)
            cnt = ema_count * decay + (1 - decay) * counts
            new_cnt = (cnt + epsilon) / (b + K * epsilon) * b
            dw = torch.zeros_like(codebook).index_add_(0, idx, flat.detach().to(codebook.dtype))   # no-op in fp32 (bench: autocast leg)
            new_w = ema_weight * decay + (1 - decay) * dw
            new_cb = new_w / new_cnt.unsqueeze(1)
    e_loss = beta * F.mse_loss(q.detach(), flat)
    q_st = flat + (q - flat).detach()
    return _unflatten(q_st, z.shape), idx.reshape(z.shape[0], -1).detach(), e_loss, new_cb, new_cnt, new_w


def entropy_term(affinity: Tensor, temperature: float, loss_type: str = 'softmax') -> Tensor:
    """vector_quantizers.py:296-328: mean sample entropy minus entropy of the batch-mean distribution."""
    n_classes = affinity.shape[-1]
    a = affinity / temperature
    probs = F.softmax(a, dim=-1)
    if loss_type == 'softmax':
        target = probs
    elif loss_type == 'argmax':
        hard = F.one_hot(torch.argmax(a, dim=-1), n_classes).to(torch.int64)
        hard = hard.to(torch.argmax(a, dim=-1))          # the reference casts one-hots to the index dtype
        target = probs - (probs - hard).detach()
    else:
        raise ValueError('Entropy loss {} not supported'.format(loss_type))
    avg = target.mean(dim=0)
    avg_entropy = -torch.sum(avg * torch.log(avg + 1e-5))
    logp = F.log_softmax(a + 1e-5, dim=-1)
    sample_entropy = torch.mean(-torch.sum(target * logp, dim=-1))
    return sample_entropy - avg_entropy


def vq_entropy(z: Tensor, codebook: Tensor, beta: float, ent_ratio: float, ent_temperature: float,
               ent_loss_type: str = 'softmax'):
    """vector_quantizers.py:290-356."""
    flat = _flatten_latents(z)
    d = l2_distances(flat, codebook, order='entropy')
    idx = torch.argmin(d, dim=1)
    q = _unflatten(F.embedding(idx, codebook), z.shape)
    e_loss = torch.mean((q.detach() - z) ** 2) * beta
    q_loss = torch.mean((q - z.detach()) ** 2)
    ent = entropy_term(-d, ent_temperature, ent_loss_type) * ent_ratio
    q_st = z + (q - z).detach()
    return q_st, idx.reshape(z.shape[0], -1).detach(), e_loss + q_loss + ent


def vq_gumbel(x: Tensor, codebook: Tensor, proj_w: Tensor, proj_b: Tensor, temp: float, kl_cost: float,
              hard: bool, exp_noise: Tensor):
    """vector_quantizers.py:223-245 with the Gumbel noise made explicit (defect B6: noise is always added).
    F.gumbel_softmax draws E ~ Exp(1) and uses g = -log(E); ``exp_noise`` is that E tensor (same shape as
    the logits) so the CUDA kernel and the oracle consume identical noise.  idx keeps the reference's
    (B,H,W) shape (defect B5)."""
    logits = F.conv2d(x, proj_w, proj_b)
    g = -torch.log(exp_noise)
    y_soft = F.softmax((logits + g) / temp, dim=1)
    if hard:
        top = y_soft.argmax(dim=1, keepdim=True)
        y_hard = torch.zeros_like(logits).scatter_(1, top, 1.0)
        y = y_hard - y_soft.detach() + y_soft
    else:
        y = y_soft
    q = torch.einsum('bnhw,nd->bdhw', y, codebook)
    qy = F.softmax(logits, dim=1)
    K = codebook.shape[0]
    kl = kl_cost * torch.sum(qy * torch.log(qy * K + 1e-10), dim=1).mean()
    return q, y.argmax(dim=1).detach(), kl


def codebook_usage(index_count: Tensor):
    """abstract_modules/base_quantizer.py:63-79."""
    p = index_count / torch.sum(index_count)
    perplexity = torch.exp(-torch.sum(p * torch.log(p + 1e-10), dim=-1)).sum().item()
    used = torch.count_nonzero(p).item() * 100 / index_count.shape[0]
    return p, perplexity, used


# --------------------------------------------------------------------------------------
# optimizer grouping + AdamW  (vqvae/model.py:372-440)
# --------------------------------------------------------------------------------------
def adamw_groups(names_enc: Sequence[str], names_dec: Sequence[str], names_q: Sequence[str],
                 replicate_name_collision: bool = True):
    """model.py:372-440.  The reference builds ``param_dict`` from names RELATIVE to each sub-module, so an
    encoder and a decoder parameter with the same relative name collide and the encoder tensor is dropped
    (defect B2).  Returns (decay_names, no_decay_names) as FULL names ('encoder.conv_in.weight'); with
    ``replicate_name_collision`` the colliding encoder tensors are omitted exactly as in the reference.
    Decay: Conv2d weights.  No decay: every bias, Embedding weight, GroupNorm weight."""
    def is_decay(rel: str) -> bool:
        if rel.endswith('bias'):
            return False
        leaf = rel.rsplit('.', 2)[-2] if '.' in rel else ''
        if leaf.startswith('norm') or rel == 'codebook.weight':
            return False
        return True

    chosen: Dict[str, str] = {}
    for prefix, names in (('encoder.', names_enc), ('decoder.', names_dec), ('quantizer.', names_q)):
        for rel in names:
            if replicate_name_collision:
                chosen[rel] = prefix + rel          # later sub-module overwrites the earlier one
            else:
                chosen[prefix + rel] = prefix + rel
    decay, no_decay = [], []
    for rel in sorted(chosen):
        rel_name = rel if replicate_name_collision else rel.split('.', 1)[1]
        (decay if is_decay(rel_name) else no_decay).append(chosen[rel])
    return decay, no_decay


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float, beta2: float,
               eps: float, weight_decay: float) -> None:
    """torch.optim.AdamW single-tensor semantics (decoupled decay, bias-corrected), in place."""
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# schedules (scheduling_utils.schedulers_cpp -- external, un-vendored, no version pin: PARITY UNPINNED;
# semantics inferred from the call sites vqvae/model.py:175-200,210-224)
# --------------------------------------------------------------------------------------
def linear_schedule(step: int, start_step: int, stop_step: int, start_value: float, stop_value: float) -> float:
    if step <= start_step:
        return start_value
    if step >= stop_step:
        return stop_value
    t = (step - start_step) / float(stop_step - start_step)
    return start_value + (stop_value - start_value) * t


def cosine_schedule(step: int, start_step: int, stop_step: int, start_value: float, stop_value: float) -> float:
    if step <= start_step:
        return start_value
    if step >= stop_step:
        return stop_value
    t = (step - start_step) / float(stop_step - start_step)
    return stop_value + (start_value - stop_value) * 0.5 * (1.0 + math.cos(math.pi * t))


def linear_cosine_schedule(step: int, start_step: int, stop_step: int, start_value: float, stop_value: float,
                           th_step: int) -> float:
    if step < th_step:
        return linear_schedule(step, start_step, th_step, 1e-20, start_value)
    return cosine_schedule(step, th_step, stop_step, start_value, stop_value)


# --------------------------------------------------------------------------------------
# one training step, plain VQ-VAE branch (vqvae/model.py:232-295 branch C, + optimizer)
# --------------------------------------------------------------------------------------
def ssim_torchmetrics(preds: Tensor, target: Tensor, kernel_size: int = 11, sigma: float = 1.5, k1: float = 0.01,
                      k2: float = 0.03) -> Tensor:
    """Per-image SSIM as the reference's evaluation computes it (vqvae/model.py:495, 529-530: torchmetrics
    StructuralSimilarityIndexMeasure() with its defaults).  PARITY UNPINNED: torchmetrics is a third-party dependency that is
    absent from this image; this restates its published algorithm (functional/image/ssim.py, v1.x): reflect-pad by
    (kernel_size - 1) / 2, depth-wise Gaussian filtering of (x, y, x^2, y^2, xy), variances clamped at 0, the padded border
    cropped again, mean over (C, H', W') per image; data_range=None -> max of the two batch ranges."""
    c = preds.shape[1]
    data_range = torch.maximum(preds.max() - preds.min(), target.max() - target.min())
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=preds.dtype)
    g = torch.exp(-((dist / sigma) ** 2) / 2)
    g = (g / g.sum()).unsqueeze(0)
    kernel = (g.t() @ g).expand(c, 1, kernel_size, kernel_size)
    pad = (kernel_size - 1) // 2
    p = F.pad(preds, (pad, pad, pad, pad), mode='reflect')
    t = F.pad(target, (pad, pad, pad, pad), mode='reflect')
    stack = torch.cat((p, t, p * p, t * t, p * t))
    out = F.conv2d(stack, kernel, groups=c).split(preds.shape[0])
    mu_p2, mu_t2, mu_pt = out[0] ** 2, out[1] ** 2, out[0] * out[1]
    s_p = torch.clamp(out[2] - mu_p2, min=0.0)
    s_t = torch.clamp(out[3] - mu_t2, min=0.0)
    s_pt = out[4] - mu_pt
    upper, lower = 2 * s_pt + c2, s_p + s_t + c2
    full = ((2 * mu_pt + c1) * upper) / ((mu_p2 + mu_t2 + c1) * lower)
    return full[..., pad:-pad, pad:-pad].reshape(preds.shape[0], -1).mean(-1)


def normalize_images(images01: Tensor) -> Tensor:
    """abstract_modules/base_autoencoder.py:41-50 without the kornia augmentation: clamp to [0,1], (x-.5)/.5."""
    return (torch.clamp(images01, 0.0, 1.0) - 0.5) / 0.5


def forward_vqvae(sd: SD, images: Tensor, cfg: dict, training: bool = True, exp_noise: Optional[Tensor] = None,
                  batch_for_smoothing: Optional[int] = None):
    """model.py:151-161.  ``cfg`` = {'num_res_blocks', 'channel_multipliers', 'quantizer': {'type', ...params}}.
    Returns dict(z, quantized, idx, q_loss, recon [, new_codebook, new_ema_count, new_ema_weight])."""
    nrb, mult, q = cfg['num_res_blocks'], cfg['channel_multipliers'], cfg['quantizer']
    z = encoder_forward(sd, images, nrb, mult)
    out = {'z': z}
    cb = sd['quantizer.codebook.weight']
    if q['type'] == 'standard':
        quant, idx, ql = vq_standard(z, cb, q['commitment_cost'])
    elif q['type'] == 'ema':
        quant, idx, ql, ncb, ncnt, nw = vq_ema(z, cb, sd['quantizer.ema_count'], sd['quantizer.ema_weight'],
                                               q['commitment_cost'], q['decay'], q['epsilon'], training,
                                               batch_for_smoothing)
        out.update(new_codebook=ncb, new_ema_count=ncnt, new_ema_weight=nw)
    elif q['type'] == 'entropy':
        quant, idx, ql = vq_entropy(z, cb, q['commitment_cost'], q['ent_loss_ratio'], q['ent_temperature'],
                                    q.get('ent_loss_type', 'softmax'))
    elif q['type'] == 'gumbel':
        hard = q['straight_through'] if training else True
        quant, idx, ql = vq_gumbel(z, cb, sd['quantizer.x_to_logits.weight'], sd['quantizer.x_to_logits.bias'],
                                   q['temp'], q['kl_cost'], hard, exp_noise)
    else:
        raise ValueError(f'unrecognized quantizer: {q["type"]}')
    recon = decoder_forward(sd, quant, nrb, mult)
    out.update(quantized=quant, idx=idx, q_loss=ql, recon=recon)
    return out


def train_step_mse(sd: SD, images: Tensor, cfg: dict, exp_noise: Optional[Tensor] = None):
    """Forward + backward of branch C (model.py:272-275): loss = q_loss + mse(recon, images).
    ``sd`` tensors that require grad get their ``.grad`` populated.  Returns the forward dict + losses."""
    out = forward_vqvae(sd, images, cfg, training=True, exp_noise=exp_noise)
    l2 = F.mse_loss(out['recon'], images)
    loss = out['q_loss'] + l2
    loss.backward()
    out.update(l2_loss=l2, loss=loss)
    return out
