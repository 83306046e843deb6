"""StyleGAN2-ADA discriminator ('resnet' architecture) on the libvqgan_b200 kernels.

Reference: vqvae/modules/loss/stylegan2_discriminator/discriminator.py (FullyConnectedLayer :90-121, Conv2dLayer :126-174,
DiscriminatorBlock :179-265, MinibatchStdLayer :270-293, DiscriminatorEpilogue :298-354, Discriminator :359-414) and its
ops (conv2d_resample.py:59-154, upfirdn2d.py, bias_act.py).  Same module tree, parameter / buffer names, shapes and
torch.randn consumption order, so reference checkpoints load and seeded initialisation matches.  Equalised-lr weight
gains, bias, leaky-ReLU(0.2) * gain and the residual add are epilogues of the implicit-GEMM conv; FIR low-pass
resampling, minibatch-stddev and the NCHW flatten are their own HBM-bound kernels."""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from ... import ops, ops_gan
from ...lib import ACT_LRELU, ACT_NONE

_SQRT2 = math.sqrt(2.0)


def setup_filter(f=(1, 3, 3, 1)) -> torch.Tensor:
    """upfirdn2d.setup_filter for the separable 4-tap case: outer(f, f) / sum  (upfirdn2d.py:72-121)"""
    f = torch.as_tensor(f, dtype=torch.float32)
    f = f.ger(f)
    return f / f.sum()


class FullyConnectedLayer(nn.Module):
    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.activation = activation
        self.weight = nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier
        if self.bias_gain != 1:
            raise NotImplementedError('lr_multiplier != 1 is only used by the (unused) mapping network')

    def forward(self, x):
        """x: channels-last [N, in_features, 1, 1] -> [N, out_features, 1, 1] fp32"""
        w = self.weight.reshape(self.weight.shape[0], self.weight.shape[1], 1, 1)
        if self.activation == 'linear':
            return ops.conv2d(x, w, self.bias, out_dtype=torch.float32, w_scale=float(self.weight_gain))
        if self.activation != 'lrelu':
            raise NotImplementedError(self.activation)
        return ops.conv2d(x, w, self.bias, act=ACT_LRELU, alpha=0.2, gain=_SQRT2, out_dtype=torch.float32,
                          w_scale=float(self.weight_gain))


class Conv2dLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation='linear', up=1, down=1,
                 resample_filter=(1, 3, 3, 1), conv_clamp=None, channels_last=False, trainable=True):
        super().__init__()
        if up != 1 or conv_clamp is not None or not trainable:
            raise NotImplementedError('only the configurations the VQGAN discriminator uses are built')
        self.activation = activation
        self.up, self.down = up, down
        self.register_buffer('resample_filter', setup_filter(resample_filter))
        self.kernel_size = kernel_size
        self.padding = kernel_size // 2
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))
        self.act_gain = _SQRT2 if activation == 'lrelu' else 1.0
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = nn.Parameter(torch.zeros([out_channels])) if bias else None

    def forward(self, x, gain=1, residual=None):
        act_gain = float(self.act_gain * gain)
        w_scale = float(self.weight_gain)
        if self.activation == 'lrelu':
            kw = dict(act=ACT_LRELU, alpha=0.2, gain=act_gain)
        elif self.activation == 'linear':
            if self.bias is not None and act_gain != 1.0:
                raise NotImplementedError('linear layer with bias and gain')
            w_scale *= act_gain                        # (conv * g) == conv with (w * g): no bias on this path
            kw = dict(act=ACT_NONE)
        else:
            raise NotImplementedError(self.activation)
        if self.down == 1:
            co, ci = self.weight.shape[0], self.weight.shape[1]
            if ops.get_precision().name == 'fast' and ci > 64 and ci % 64 != 0 and co % 128 == 0 and x.shape[1] == ci:
                # the epilogue's 513 -> 512 conv (512 + the minibatch-stddev channel): zero-pad the input channels to the next
                # multiple of 64 so that forward, dgrad and wgrad ride the tensor cores (tiny tensors: N x 576 x 4 x 4 and a
                # 10 MB weight copy; the padding ops are layout plumbing and autograd slices the gradients back)
                cpad = (ci + 63) // 64 * 64 - ci
                xp = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, cpad)).contiguous(memory_format=torch.channels_last)
                wp = torch.nn.functional.pad(self.weight, (0, 0, 0, 0, 0, cpad))
                return ops.conv2d(xp, wp, self.bias, residual, pad=self.padding, w_scale=w_scale, **kw)
            return ops.conv2d(x, self.weight, self.bias, residual, pad=self.padding, w_scale=w_scale, **kw)
        if self.down != 2:
            raise NotImplementedError('down must be 1 or 2')
        if self.kernel_size == 1:
            # conv2d_resample.py:107-110: FIR (pad 1) + keep every 2nd sample, then the 1x1 conv
            x = ops_gan.fir4(x, 1, 2)
            return ops.conv2d(x, self.weight, self.bias, residual, pad=0, w_scale=w_scale, **kw)
        # conv2d_resample.py:119-122: pad k//2 + 1 each side, FIR, then a stride-2 conv without padding
        if (self.kernel_size == 3 and residual is None and not ops_gan.in_second_order() and _S2D_ROUTE
                and ops_gan.down2_conv3x3_supported(x, self.weight)):
            # space-to-depth route: FIR written in 2x2 space-to-depth layout, then a 2x2-tap tensor-core convolution over 4C channels
            act = kw['act']
            return ops_gan.down2_conv3x3(x, self.weight, self.bias, act, kw.get('alpha', 0.0), kw.get('gain', 1.0), w_scale)
        x = ops_gan.fir4(x, self.padding + 1, 1)
        co, ci = self.weight.shape[0], self.weight.shape[1]
        if (ops.get_precision().name == 'fast' and self.kernel_size == 3 and residual is None and ci % 64 == 0 and co % 128 == 0
                and x.shape[2] >= 9 and x.shape[3] >= 9):
            # tensor-core route: the same convolution at stride 1 ('same' padding, full resolution) keeps every 2nd output:
            # valid_out[i] = same_out[i+1], strided_out[o] = valid_out[2o] = same_out[2o+1]
            full = ops.conv2d(x, self.weight, self.bias, None, pad=1, stride=1, w_scale=w_scale, **kw)
            return ops_gan.decimate2(full, (x.shape[2] - 3) // 2 + 1, (x.shape[3] - 3) // 2 + 1, 1)
        return ops.conv2d(x, self.weight, self.bias, residual, pad=0, stride=2, w_scale=w_scale, **kw)


import os as _os
_S2D_ROUTE = _os.environ.get('VQB_D_S2D', '1') != '0'      # A/B switch: 0 = full-resolution stride-1 conv + decimation


class DiscriminatorBlock(nn.Module):
    def __init__(self, in_channels, tmp_channels, out_channels, resolution, img_channels, first_layer_idx,
                 architecture='resnet', activation='lrelu', resample_filter=(1, 3, 3, 1), conv_clamp=None, use_fp16=False,
                 fp16_channels_last=False, freeze_layers=0):
        assert in_channels in [0, tmp_channels]
        if architecture != 'resnet':
            raise NotImplementedError("only the 'resnet' architecture (the reference default) is built")
        super().__init__()
        self.in_channels = in_channels
        self.resolution = resolution
        self.img_channels = img_channels
        self.first_layer_idx = first_layer_idx
        self.architecture = architecture
        self.register_buffer('resample_filter', setup_filter(resample_filter))
        self.num_layers = 0
        if in_channels == 0:
            self.fromrgb = Conv2dLayer(img_channels, tmp_channels, kernel_size=1, activation=activation)
            self.num_layers += 1
        self.conv0 = Conv2dLayer(tmp_channels, tmp_channels, kernel_size=3, activation=activation)
        self.conv1 = Conv2dLayer(tmp_channels, out_channels, kernel_size=3, activation=activation, down=2,
                                 resample_filter=resample_filter)
        self.skip = Conv2dLayer(tmp_channels, out_channels, kernel_size=1, bias=False, down=2, resample_filter=resample_filter)
        self.num_layers += 3

    def forward(self, x, img):
        if self.in_channels == 0:
            x = self.fromrgb(img)
        h = self.conv0(x)
        h = self.conv1(h, gain=np.sqrt(0.5))
        # x = y.add_(x) fused into the epilogue of the (activation-free) skip conv: skip(x) * sqrt(.5) + conv1(...)
        return self.skip(x, gain=np.sqrt(0.5), residual=h), None


class MinibatchStdLayer(nn.Module):
    def __init__(self, group_size, num_channels=1):
        super().__init__()
        if num_channels != 1:
            raise NotImplementedError('num_channels must be 1')
        self.group_size = group_size
        self.num_channels = num_channels

    def forward(self, x):
        return ops_gan.mbstd(x, self.group_size)


class DiscriminatorEpilogue(nn.Module):
    def __init__(self, in_channels, cmap_dim, resolution, img_channels, architecture='resnet', mbstd_group_size=4,
                 mbstd_num_channels=1, activation='lrelu', conv_clamp=None):
        super().__init__()
        if cmap_dim != 0:
            raise NotImplementedError('conditional discriminator')
        self.in_channels = in_channels
        self.cmap_dim = cmap_dim
        self.resolution = resolution
        self.mbstd = MinibatchStdLayer(group_size=mbstd_group_size, num_channels=mbstd_num_channels) if mbstd_num_channels > 0 else None
        self.conv = Conv2dLayer(in_channels + mbstd_num_channels, in_channels, kernel_size=3, activation=activation)
        self.fc = FullyConnectedLayer(in_channels * (resolution ** 2), in_channels, activation=activation)
        self.out = FullyConnectedLayer(in_channels, 1)

    def forward(self, x, img=None, cmap=None):
        if self.mbstd is not None:
            x = self.mbstd(x)
        x = self.conv(x)
        x = self.fc(ops_gan.flatten_nchw(x))
        x = self.out(x)
        return x.reshape(x.shape[0], -1)


class Discriminator(nn.Module):
    def __init__(self, img_resolution, c_dim=0, img_channels=3, architecture='resnet', channel_base=32768, channel_max=512,
                 num_fp16_res=0, conv_clamp=None, cmap_dim=None):
        super().__init__()
        if c_dim != 0:
            raise NotImplementedError('conditional discriminator')
        self.c_dim = c_dim
        self.img_resolution = img_resolution
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.img_channels = img_channels
        self.block_resolutions = [2 ** i for i in range(self.img_resolution_log2, 2, -1)]
        channels_dict = {res: min(channel_base // res, channel_max) for res in self.block_resolutions + [4]}
        cur_layer_idx = 0
        for res in self.block_resolutions:
            in_channels = channels_dict[res] if res < img_resolution else 0
            block = DiscriminatorBlock(in_channels, channels_dict[res], channels_dict[res // 2], resolution=res,
                                       img_channels=img_channels, first_layer_idx=cur_layer_idx, architecture=architecture)
            setattr(self, f'b{res}', block)
            cur_layer_idx += block.num_layers
        self.b4 = DiscriminatorEpilogue(channels_dict[4], cmap_dim=0, resolution=4, img_channels=img_channels,
                                        architecture=architecture)

    def forward(self, img):
        """img [B,3,R,R] in [-1,1] -> logits [B,1] fp32"""
        x = None
        img = ops.as_nhwc(img)
        for res in self.block_resolutions:
            x, _ = getattr(self, f'b{res}')(x, img)
        return self.b4(x)
