"""LPIPS perceptual loss on the libvqgan_b200 kernels (reference: vqvae/modules/loss/lpips_pytorch/{modules/lpips.py,
modules/networks.py, modules/utils.py}).  Same module tree / state_dict keys as the reference
(`net.mean`, `net.std`, `net.layers.N.{weight,bias}`, `lin.N.1.weight`): frozen VGG16 `features[0:30]` as implicit-GEMM
convolutions with fused bias+ReLU epilogues, 2x2 max-pools, and ONE fused kernel per feature tap for
channel-unit-normalise -> squared difference -> 1x1 lin -> spatial / batch mean."""
from __future__ import annotations

from typing import Sequence

import torch
from torch import nn

from ... import ops, ops_gan
from ...lib import ACT_RELU
from ..autoencoder import Conv2d

VGG16_CFG = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512)   # features[0:30]


class _FusedReLU(nn.Identity):
    """placeholder keeping torchvision's layer numbering; the ReLU runs in the preceding conv's epilogue"""


class _MaxPool(nn.Module):
    def forward(self, x):
        return ops_gan.max_pool2(x)


class _MaxPool3s2(nn.Module):
    def forward(self, x):
        return ops_gan.max_pool3s2(x)


class LinLayers(nn.ModuleList):
    """networks.py:24-34 -- frozen 1x1 convs (nc -> 1, no bias); applied inside the fused tap kernel"""

    def __init__(self, n_channels_list: Sequence[int]):
        super().__init__([nn.Sequential(nn.Identity(), nn.Conv2d(nc, 1, 1, 1, 0, bias=False)) for nc in n_channels_list])
        for p in self.parameters():
            p.requires_grad = False


class VGG16(nn.Module):
    """networks.py:37-64,89-97: z-score, then features with taps after relu1_2, relu2_2, relu3_3, relu4_3, relu5_3."""

    def __init__(self):
        super().__init__()
        self.register_buffer('mean', torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('std', torch.Tensor([.458, .448, .450])[None, :, None, None])
        layers, cin = [], 3
        for v in VGG16_CFG:
            if v == 'M':
                layers.append(_MaxPool())
            else:
                layers += [Conv2d(cin, v, kernel_size=3, padding=1), _FusedReLU()]
                cin = v
        self.layers = nn.Sequential(*layers)
        self.target_layers = [4, 9, 16, 23, 30]
        self.n_channels_list = [64, 128, 256, 512, 512]
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, x: torch.Tensor):
        scale = (1.0 / self.std).reshape(-1).float().contiguous()
        shift = (-self.mean / self.std).reshape(-1).float().contiguous()
        x = ops_gan.channel_affine(x, scale, shift, torch.float32)
        out = []
        for i, layer in enumerate(self.layers, 1):
            if isinstance(layer, Conv2d):
                x = layer(x, act=ACT_RELU)
            elif isinstance(layer, _MaxPool):
                x = layer(x)
            if i in self.target_layers:
                out.append(x)
        return out


class _BaseNet(nn.Module):
    """networks.py:37-64: z-score with the LPIPS constants, then the trunk with normalised taps after `target_layers`."""

    def __init__(self):
        super().__init__()
        self.register_buffer('mean', torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('std', torch.Tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, x: torch.Tensor):
        scale = (1.0 / self.std).reshape(-1).float().contiguous()
        shift = (-self.mean / self.std).reshape(-1).float().contiguous()
        x = ops_gan.channel_affine(x, scale, shift, torch.float32)
        out = []
        for i, layer in enumerate(self.layers, 1):
            if isinstance(layer, Conv2d):
                x = layer(x, act=ACT_RELU)
            elif isinstance(layer, (_MaxPool, _MaxPool3s2)):
                x = layer(x)
            if i in self.target_layers:
                out.append(x)
            if len(out) == len(self.target_layers):
                break
        return out


class AlexNet(_BaseNet):
    """networks.py:78-86: torchvision alexnet.features (conv 11x11/4, pool 3/2, conv 5x5, pool 3/2, three 3x3 convs) with taps
    after every ReLU (layers 2, 5, 8, 10, 12; 64, 192, 384, 256, 256 channels) -- the trunk of the VQLPIPS ablation loss
    (loss.py:182).  Same layer numbering, hence the same state_dict keys, as torchvision."""

    def __init__(self):
        super().__init__()
        self.layers = nn.Sequential(
            Conv2d(3, 64, kernel_size=11, stride=4, padding=2), _FusedReLU(), _MaxPool3s2(),
            Conv2d(64, 192, kernel_size=5, padding=2), _FusedReLU(), _MaxPool3s2(),
            Conv2d(192, 384, kernel_size=3, padding=1), _FusedReLU(),
            Conv2d(384, 256, kernel_size=3, padding=1), _FusedReLU(),
            Conv2d(256, 256, kernel_size=3, padding=1), _FusedReLU(), _MaxPool3s2())
        self.target_layers = [2, 5, 8, 10, 12]
        self.n_channels_list = [64, 192, 384, 256, 256]
        for p in self.parameters():
            p.requires_grad = False


class LPIPS(nn.Module):
    """LPIPS(net_type)(x, y) -> scalar (lpips.py:18-38).  `pretrained=True` loads torchvision's VGG16 weights and the
    LPIPS v0.1 lin weights exactly as the reference does (both need a network connection or a populated torch hub cache)
    and raises if they are unavailable; `pretrained=False` leaves the seeded random initialisation (parity tests)."""

    def __init__(self, net_type: str = 'alex', version: str = '0.1', pretrained: bool = True):
        assert version in ['0.1'], 'v0.1 is only supported now'
        super().__init__()
        if net_type == 'vgg':
            self.net = VGG16()
        elif net_type == 'alex':
            self.net = AlexNet()
        else:
            raise NotImplementedError("choose net_type from [alex, vgg] (the reference's 'squeeze' trunk is used nowhere in it)")
        self.lin = LinLayers(self.net.n_channels_list)
        if pretrained:
            self._load_pretrained(net_type, version)
        for p_ in self.net.parameters():          # frozen trunk: its kernel-layout weight copies are packed once, not every step
            p_._vqb_static = True

    def _load_pretrained(self, net_type: str, version: str) -> None:
        from torchvision import models
        if net_type == 'vgg':
            tv = models.vgg16(weights=models.VGG16_Weights.DEFAULT).features        # networks.py:93
            self.net.layers.load_state_dict({k: v for k, v in tv.state_dict().items() if int(k.split('.')[0]) < 30})
        else:
            tv = models.alexnet(weights=models.AlexNet_Weights.DEFAULT).features    # networks.py:82 (alexnet(True))
            self.net.layers.load_state_dict(tv.state_dict())
        url = ('https://raw.githubusercontent.com/richzhang/PerceptualSimilarity/' + f'master/lpips/weights/v{version}/{net_type}.pth')
        sd = torch.hub.load_state_dict_from_url(url, progress=False, map_location='cpu')          # utils.py:11-20
        self.lin.load_state_dict({k.replace('lin', '').replace('model.', ''): v for k, v in sd.items()})

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            feat_x = self.net(x)                                   # target images: no gradient needed
        feat_y = self.net(y)
        total = None
        for fx, fy, lin in zip(feat_x, feat_y, self.lin):
            t = ops_gan.lpips_tap(fx, fy, lin[1].weight)
            total = t if total is None else total + t
        return total
