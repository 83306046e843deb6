"""Encoder / Decoder of the VQ-VAE -- same classes, constructor signatures, parameter names/shapes and RNG
consumption order as the reference (vqvae/modules/autoencoder.py), with every forward/backward op executed by the
libvqgan_b200 CUDA kernels (implicit-GEMM convolutions with fused bias / residual / tanh epilogues, fused
GroupNorm+SiLU, NHWC activations)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..lib import ACT_NONE, ACT_SILU, ACT_TANH


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameters (same init, same state_dict keys); forward = vqb_conv2d_fwd with fused epilogue."""

    def forward(self, x, residual=None, act=ACT_NONE, out_dtype=None, gn_groups=0):
        """gn_groups: number of groups of the GroupNorm that consumes this output (its statistics then come out of the
        convolution's epilogue, ops.conv2d)"""
        pad = self.padding
        if isinstance(pad, str):
            if pad != 'same':
                raise ValueError(f'unsupported padding {pad}')
            pad = self.kernel_size[0] // 2
        else:
            pad = pad[0]
        return ops.conv2d(x, self.weight, self.bias, residual, pad=pad, stride=self.stride[0], act=act, out_dtype=out_dtype,
                          gn_groups=gn_groups)


class GroupNorm(nn.Module):
    """reference: autoencoder.py:7-39 -- affine parameters shaped (1,C,1,1), eps 1e-6, UNBIASED variance."""

    def __init__(self, num_groups: int, num_channels: int, eps: float = 1e-6):
        super().__init__()
        if num_channels % num_groups != 0:
            raise ValueError('num_channels must be divisible by num_groups')
        self.num_groups = num_groups
        self.num_channels = num_channels
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(1, num_channels, 1, 1))
        self.bias = nn.Parameter(torch.zeros(1, num_channels, 1, 1))

    def forward(self, x: torch.Tensor, act: int = ACT_NONE, want_skip: bool = False):
        return ops.group_norm_act(x, self.weight, self.bias, self.num_groups, self.eps, act, want_skip)


class ResBlock(nn.Module):
    """reference: autoencoder.py:42-77.  GN+SiLU are one kernel pair, `x + h` is the epilogue of conv2."""

    def __init__(self, in_channels: int, out_channels: int = None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        if self.in_channels != self.out_channels:
            self.conv_shortcut = Conv2d(self.in_channels, self.out_channels, kernel_size=1, stride=1, padding='same', bias=False)
        self.norm1 = GroupNorm(32, self.in_channels)
        self.conv1 = Conv2d(self.in_channels, self.out_channels, kernel_size=3, stride=1, padding='same', bias=False)
        self.norm2 = GroupNorm(32, self.out_channels)
        self.conv2 = Conv2d(self.out_channels, self.out_channels, kernel_size=3, stride=1, padding='same', bias=False)

    def forward(self, x):
        if self.in_channels != self.out_channels:
            h = self.norm1(x, act=ACT_SILU)
            x = self.conv_shortcut(x)
        else:
            # the skip connection goes through norm1's identity output: its gradient is added inside the GN backward kernel
            h, x = self.norm1(x, act=ACT_SILU, want_skip=True)
        h = self.conv1(h, gn_groups=self.norm2.num_groups)          # norm2's statistics come out of conv1's epilogue
        h = self.norm2(h, act=ACT_SILU)
        # the block output is (almost always) the input of the next GroupNorm: next block's norm1 or the final norm
        return self.conv2(h, residual=x, gn_groups=self.norm2.num_groups)


class Downsample(nn.Module):
    """reference: autoencoder.py:80-91 (avg_pool2d 2x2 stride 2)."""

    def __init__(self, kernel_size: int = 2, stride: int = 2, padding: int = 0):
        super().__init__()
        if (kernel_size, stride, padding) != (2, 2, 0):
            raise ValueError('only the 2x2/stride-2 average pool used by the reference is implemented')

    def forward(self, x):
        return ops.avg_pool2(x)


class Upsample(nn.Module):
    """reference: autoencoder.py:94-106 (nearest-exact x2, then 3x3 conv with bias)."""

    def __init__(self, channels: int, scale_factor: float = 2.0, mode: str = 'nearest-exact'):
        super().__init__()
        if float(scale_factor) != 2.0 or mode not in ('nearest-exact', 'nearest'):
            raise ValueError('only the x2 nearest upsampling used by the reference is implemented')
        self.conv = Conv2d(channels, channels, kernel_size=3, padding='same')

    def forward(self, x):
        return self.conv(ops.upsample2(x), gn_groups=32)            # feeds the next ResBlock's norm1 / the decoder's final norm


class Encoder(nn.Module):
    """reference: autoencoder.py:109-143."""

    def __init__(self, channels: int, num_res_blocks: int, channel_multipliers: tuple, embedding_dim: int):
        super().__init__()
        self.conv_in = Conv2d(3, channels, kernel_size=3, padding='same', bias=False)
        blocks = []
        ch_in = channels
        for i in range(len(channel_multipliers)):
            ch_out = channels * channel_multipliers[i]
            for _ in range(num_res_blocks):
                blocks.append(ResBlock(ch_in, ch_out))
                ch_in = ch_out
            blocks.append(Downsample())
        self.blocks = nn.Sequential(*blocks)
        self.final_residual = nn.Sequential(*[ResBlock(ch_in) for _ in range(num_res_blocks)])
        self.norm = GroupNorm(32, ch_in)
        self.conv_out = Conv2d(ch_in, embedding_dim, kernel_size=1, padding='same')

    def forward(self, x):
        """x: [B,3,H,W] in [-1,1] (NCHW or channels-last) -> z [B,embedding_dim,H/2^L,W/2^L] fp32, channels-last."""
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = self.conv_in(ops.as_nhwc(x), gn_groups=32)
        x = self.blocks(x)
        x = self.final_residual(x)
        x = self.norm(x, act=ACT_SILU)
        return self.conv_out(x, out_dtype=torch.float32)


class Decoder(nn.Module):
    """reference: autoencoder.py:146-180."""

    def __init__(self, channels: int, num_res_blocks: int, channel_multipliers: tuple, embedding_dim: int):
        super().__init__()
        ch_in = channels * channel_multipliers[-1]
        self.conv_in = Conv2d(embedding_dim, ch_in, kernel_size=3, padding='same')
        self.initial_residual = nn.Sequential(*[ResBlock(ch_in) for _ in range(num_res_blocks)])
        blocks = []
        for i in reversed(range(len(channel_multipliers))):
            ch_out = channels * channel_multipliers[i - 1] if i > 0 else channels
            for _ in range(num_res_blocks):
                blocks.append(ResBlock(ch_in, ch_out))
                ch_in = ch_out
            blocks.append(Upsample(ch_out))
        self.blocks = nn.Sequential(*blocks)
        self.norm = GroupNorm(32, channels)
        self.conv_out = Conv2d(channels, 3, kernel_size=3, padding='same')

    def forward(self, x):
        """x: [B,embedding_dim,h,w] -> reconstruction [B,3,H,W] fp32 in [-1,1] (tanh fused in conv_out's epilogue)."""
        x = self.conv_in(ops.as_nhwc(x), gn_groups=32)
        x = self.initial_residual(x)
        x = self.blocks(x)
        x = self.norm(x, act=ACT_SILU)
        return self.conv_out(x, act=ACT_TANH, out_dtype=torch.float32)
