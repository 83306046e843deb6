"""Training-time input pipeline: clamp -> random resized crop -> horizontal flip -> normalise, ONE kernel
(vqb_crop_flip_normalize) on the batch as the loader delivers it (NCHW fp32 / fp16 in [0,1] or uint8 in [0,255]).

Reference: BaseVQVAE.__init__ / preprocess_batch (vqvae/modules/abstract_modules/base_autoencoder.py:17-50), which chains
kornia's RandomResizedCrop((S,S), scale=(0.7,1.0), ratio=(1,1)) and RandomHorizontalFlip() (per-sample parameters,
bilinear resampling with align_corners=True, integer crop boxes) and then Normalize(0.5, 0.5).

PARITY UNPINNED for the random parameter stream: kornia is an un-vendored, un-pinned dependency that is absent here
(SURVEY.md 8c), so the box / flip sampling restates its documented behaviour (area ~ U(scale)*H*W, side = round(sqrt(area)),
corner ~ floor(U(0, W - side + 1)), flip with p = 0.5) with torch's device RNG.  Given the SAME boxes and flips the kernel
is checked against torch's own crop + F.interpolate(bilinear, align_corners=True) + flip + normalise (tests/test_augment_gpu.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .lib import call, dt, ptr, stream
from .ops import empty_nhwc

_IN_DTYPES = {torch.float32: 0, torch.float16: 1, torch.uint8: 2}


def crop_flip_normalize(images: torch.Tensor, boxes: torch.Tensor, flip: Optional[torch.Tensor], out_size: Tuple[int, int],
                        out_dtype: torch.dtype = torch.float32, mean: float = 0.5, std: float = 0.5) -> torch.Tensor:
    """images [N,C,H,W] (NCHW-contiguous; fp32/fp16 in [0,1] or uint8) -> channels-last [N,C,OH,OW] in [-1,1].
    boxes [N,4] fp32 = (x0, y0, x1, y1) source coordinates of the first / last crop pixel; flip [N] uint8 or None."""
    if images.dtype not in _IN_DTYPES:
        raise TypeError(f'crop_flip_normalize: unsupported image dtype {images.dtype}')
    images = images.contiguous()
    n, c, h, w = images.shape
    oh, ow = out_size
    boxes = boxes.to(device=images.device, dtype=torch.float32).contiguous()
    if boxes.shape != (n, 4):
        raise ValueError(f'boxes must be [{n}, 4], got {tuple(boxes.shape)}')
    if flip is not None:
        flip = flip.to(device=images.device, dtype=torch.uint8).contiguous()
    out = empty_nhwc(n, c, oh, ow, out_dtype, images.device)
    call('vqb_crop_flip_normalize', ptr(images), _IN_DTYPES[images.dtype], ptr(out), dt(out), ptr(boxes), ptr(flip), n, c, h, w,
         oh, ow, mean, std, stream())
    return out


class RandomResizedCropFlip:
    """drop-in for the reference's `training_augmentations` (AugmentationSequential(RandomResizedCrop, RandomHorizontalFlip,
    same_on_batch=False)); `fused = True` tells preprocess_batch that the result is already normalised and channels-last."""

    fused = True

    def __init__(self, image_size: int, scale: Tuple[float, float] = (0.7, 1.0), p_flip: float = 0.5):
        self.image_size = image_size
        self.scale = scale
        self.p_flip = p_flip

    def sample(self, n: int, h: int, w: int, device, generator: Optional[torch.Generator] = None):
        """per-sample crop boxes [n,4] fp32 and flip flags [n] uint8, drawn on the device (no host synchronisation)"""
        u = torch.rand(n, 4, device=device, generator=generator)
        area = (self.scale[0] + (self.scale[1] - self.scale[0]) * u[:, 0]) * float(h * w)
        side = torch.sqrt(area).round().clamp_(1.0, float(min(h, w)))                     # ratio (1, 1): square crops
        x0 = torch.minimum(torch.floor(u[:, 1] * (w - side + 1.0)), w - side)
        y0 = torch.minimum(torch.floor(u[:, 2] * (h - side + 1.0)), h - side)
        boxes = torch.stack([x0, y0, x0 + side - 1.0, y0 + side - 1.0], dim=1)
        return boxes, (u[:, 3] < self.p_flip).to(torch.uint8)

    @torch.no_grad()
    def __call__(self, images: torch.Tensor, out_dtype: torch.dtype = torch.float32, generator=None) -> torch.Tensor:
        n, _, h, w = images.shape
        boxes, flip = self.sample(n, h, w, images.device, generator)
        return crop_flip_normalize(images, boxes, flip, (self.image_size, self.image_size), out_dtype)
