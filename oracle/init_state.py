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


LPIPS_CH = {'vgg': [64, 128, 256, 512, 512], 'alex': [64, 192, 384, 256, 256]}


def init_lpips(sd: SD, prefix: str, net_type: str) -> None:
    """RNG consumption of the reference's LPIPS(net_type) when built offline (oracle/ref_harness.patch_lpips_offline):
    torchvision trunk with weights=None (the WHOLE model is drawn, classifier included, lpips_pytorch/modules/networks.py:82,93),
    LinLayers' default Conv2d init (:24-34), then the lin weights themselves (seeded torch.rand instead of the download,
    lpips.py:29).  Keys as in the reference state_dict: net.mean, net.std, net.layers.N.{weight,bias}, lin.N.1.weight."""
    import torchvision
    tv = torchvision.models.vgg16(weights=None) if net_type == 'vgg' else torchvision.models.alexnet(weights=None)
    last = 30 if net_type == 'vgg' else 10**9
    sd[prefix + 'net.mean'] = torch.Tensor([-.030, -.088, -.188])[None, :, None, None]
    sd[prefix + 'net.std'] = torch.Tensor([.458, .448, .450])[None, :, None, None]
    for k, v in tv.features.state_dict().items():
        if net_type != 'vgg' or int(k.split('.')[0]) < last + 1:          # the reference keeps the full `features` module
            sd[prefix + 'net.layers.' + k] = v.detach().clone()
    for c in LPIPS_CH[net_type]:
        nn.Conv2d(c, 1, 1, 1, 0, bias=False)
    for i, c in enumerate(LPIPS_CH[net_type]):
        sd[prefix + f'lin.{i}.1.weight'] = torch.rand(1, c, 1, 1)


def init_discriminator(sd: SD, prefix: str, image_size: int, channel_base: int = 32768, channel_max: int = 512) -> None:
    """torch.randn draws in the constructor order of stylegan2_discriminator/discriminator.py:361-401 (blocks from the input
    resolution down to 8: fromrgb [first block only], conv0, conv1, skip; epilogue b4: conv, fc, out); biases zero;
    every Conv2dLayer / DiscriminatorBlock carries a `resample_filter` buffer = outer([1,3,3,1]) / 64."""
    import math
    f = torch.tensor([1., 3., 3., 1.])
    f = torch.outer(f, f)
    f = f / f.sum()
    log2 = int(math.log2(image_size))
    resolutions = [2 ** i for i in range(log2, 2, -1)]
    chans = {r: min(channel_base // r, channel_max) for r in resolutions + [4]}

    def conv(name, cin, cout, k, bias=True):
        sd[name + '.weight'] = torch.randn(cout, cin, k, k)
        if bias:
            sd[name + '.bias'] = torch.zeros(cout)
        sd[name + '.resample_filter'] = f.clone()

    for r in resolutions:
        p = f'{prefix}b{r}.'
        tmp, out = chans[r], chans[r // 2]
        sd[p + 'resample_filter'] = f.clone()
        if r == image_size:
            conv(p + 'fromrgb', 3, tmp, 1)
        conv(p + 'conv0', tmp, tmp, 3)
        conv(p + 'conv1', tmp, out, 3)
        conv(p + 'skip', tmp, out, 1, bias=False)
    c4 = chans[4]
    conv(prefix + 'b4.conv', c4 + 1, c4, 3)
    sd[prefix + 'b4.fc.weight'] = torch.randn(c4, c4 * 16)
    sd[prefix + 'b4.fc.bias'] = torch.zeros(c4)
    sd[prefix + 'b4.out.weight'] = torch.randn(1, c4)
    sd[prefix + 'b4.out.bias'] = torch.zeros(1)


def init_state(qtype: str, K: int, D: int, ch: int, nrb: int, mult: Sequence[int], seed: int = None,
               criterion: str = None, image_size: int = None) -> SD:
    """`criterion`: None (MSE, no tensors) | 'lpips' (VQLPIPS: LPIPS-AlexNet, loss.py:182) | 'gan' (VQLPIPSWithDiscriminator:
    LPIPS-VGG16 then Discriminator(image_size), loss.py:66-69) -- drawn between the decoder and init_codebook, model.py:134-149."""
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
    # --- criterion (model.py:134-145)
    if criterion == 'lpips':
        init_lpips(sd, 'criterion.perceptual_loss.', 'alex')
    elif criterion == 'gan':
        init_lpips(sd, 'criterion.perceptual_loss.', 'vgg')
        init_discriminator(sd, 'criterion.discriminator.', image_size)
    elif criterion is not None:
        raise ValueError(criterion)
    # --- init_codebook (base_quantizer.py:27-31), drawn last (model.py:148-149)
    sd['quantizer.codebook.weight'].uniform_(-1 / K, 1 / K)
    return sd


BUFFERS = ('quantizer.ema_count', 'quantizer.ema_weight')


def make_leaf(sd: SD, qtype: str) -> SD:
    """Mark trainable tensors as autograd leaves (EMA codebook is frozen, vector_quantizers.py:114)."""
    for n, t in sd.items():
        frozen = (n in BUFFERS or (qtype == 'ema' and n == 'quantizer.codebook.weight') or n.endswith('resample_filter')
                  or n.startswith('criterion.perceptual_loss.'))
        sd[n] = t.detach().clone().requires_grad_(not frozen)
    return sd
