"""Seeded parameter construction for the oracle  --  TEST INFRASTRUCTURE ONLY.

Draws every parameter/buffer in the order the reference's constructors consume the global torch RNG
(vqvae/model.py:89-149: quantizer -> Encoder -> Decoder -> init_codebook), using torch's stock layer
initialisers, and returns a flat dict keyed by the reference's state_dict names.  Checked against the
initial-weight checksums stored in tests/golden/*.npz (which come from the reference's own classes).
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
from torch import nn

SD = Dict[str, torch.Tensor]


def _conv(sd: SD, name: str, cin: int, cout: int, k: int, bias: bool) -> None:
    layer = nn.Conv2d(cin, cout, k, bias=bias)          # same default init as the reference's nn.Conv2d
    sd[name + '.weight'] = layer.weight.detach().clone()
    if bias:
        sd[name + '.bias'] = layer.bias.detach().clone()


def _norm(sd: SD, name: str, c: int) -> None:
    sd[name + '.weight'] = torch.ones(1, c, 1, 1)       # autoencoder.py:22-23
    sd[name + '.bias'] = torch.zeros(1, c, 1, 1)


def _res_block(sd: SD, p: str, cin: int, cout: int) -> None:
    """Constructor order of autoencoder.py:42-61: shortcut first, then norm1, conv1, norm2, conv2."""
    if cin != cout:
        _conv(sd, p + 'conv_shortcut', cin, cout, 1, False)
    _norm(sd, p + 'norm1', cin)
    _conv(sd, p + 'conv1', cin, cout, 3, False)
    _norm(sd, p + 'norm2', cout)
    _conv(sd, p + 'conv2', cout, cout, 3, False)


def init_state(qtype: str, K: int, D: int, ch: int, nrb: int, mult: Sequence[int], seed: int = None) -> SD:
    if seed is not None:
        torch.manual_seed(seed)
    sd: SD = {}
    # --- quantizer (abstract_modules/base_quantizer.py:21, vector_quantizers.py:114-123,216)
    sd['quantizer.codebook.weight'] = nn.Embedding(K, D).weight.detach().clone()
    if qtype == 'ema':
        sd['quantizer.ema_count'] = torch.zeros(K)
        sd['quantizer.ema_weight'] = torch.empty(K, D).uniform_(-1 / K, 1 / K)
    if qtype == 'gumbel':
        _conv(sd, 'quantizer.x_to_logits', K, K, 1, True)
    # --- encoder (autoencoder.py:110-133)
    _conv(sd, 'encoder.conv_in', 3, ch, 3, False)
    cin, i = ch, 0
    for m in mult:
        for _ in range(nrb):
            _res_block(sd, f'encoder.blocks.{i}.', cin, ch * m); cin = ch * m; i += 1
        i += 1                                           # Downsample slot (no parameters)
    for j in range(nrb):
        _res_block(sd, f'encoder.final_residual.{j}.', cin, cin)
    _norm(sd, 'encoder.norm', cin)
    _conv(sd, 'encoder.conv_out', cin, K if qtype == 'gumbel' else D, 1, True)
    # --- decoder (autoencoder.py:147-170)
    cin = ch * mult[-1]
    _conv(sd, 'decoder.conv_in', D, cin, 3, True)
    for j in range(nrb):
        _res_block(sd, f'decoder.initial_residual.{j}.', cin, cin)
    i = 0
    for lvl in reversed(range(len(mult))):
        cout = ch * mult[lvl - 1] if lvl > 0 else ch
        for _ in range(nrb):
            _res_block(sd, f'decoder.blocks.{i}.', cin, cout); cin = cout; i += 1
        _conv(sd, f'decoder.blocks.{i}.conv', cout, cout, 3, True); i += 1
    _norm(sd, 'decoder.norm', ch)
    _conv(sd, 'decoder.conv_out', ch, 3, 3, True)
    # --- init_codebook (base_quantizer.py:27-31), drawn last (model.py:148-149)
    sd['quantizer.codebook.weight'].uniform_(-1 / K, 1 / K)
    return sd


BUFFERS = ('quantizer.ema_count', 'quantizer.ema_weight')


def make_leaf(sd: SD, qtype: str) -> SD:
    """Mark trainable tensors as autograd leaves (EMA codebook is frozen, vector_quantizers.py:114)."""
    for n, t in sd.items():
        frozen = n in BUFFERS or (qtype == 'ema' and n == 'quantizer.codebook.weight')
        sd[n] = t.detach().clone().requires_grad_(not frozen)
    return sd
