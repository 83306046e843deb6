"""Pre/post-processing base class -- surface of the reference's BaseVQVAE
(vqvae/modules/abstract_modules/base_autoencoder.py:6-93)."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch

from ... import ops
from ...augment import RandomResizedCropFlip


class BaseVQVAE(ABC):

    def __init__(self, image_size: int):
        self.image_size = image_size
        # the reference applies kornia RandomResizedCrop(0.7-1, ratio 1) + RandomHorizontalFlip when training
        # (base_autoencoder.py:17-22,45-46): here one fused kernel (augment.py; random parameter stream PARITY UNPINNED,
        # kornia is absent).  Set to None to train without augmentation (parity fixtures, bench), or to any callable
        # images -> images.
        self.training_augmentations = RandomResizedCropFlip(image_size)
        self.scheduler = None
        self.train_epoch_usage_count = None
        self.val_epoch_usage_count = None

    @torch.no_grad()
    def preprocess_batch(self, images: torch.Tensor, training: bool = False) -> torch.Tensor:
        """images [B,C,H,W] in [0,1] (fp32 / fp16) or uint8 -> clamp, (training: augmentation), (x-0.5)/0.5; returned
        channels-last fp32.  One fused kernel either way (vqb_crop_flip_normalize / vqb_nchw_to_nhwc) instead of clamp +
        kornia + Normalize (base_autoencoder.py:41-50)."""
        if training and self.training_augmentations is not None:
            if getattr(self.training_augmentations, 'fused', False):
                return self.training_augmentations(images, torch.float32)
            images = self.training_augmentations(images)
        if images.dtype == torch.uint8:
            images = images.float() / 255.0
        return ops.images_to_nhwc(images, torch.float32, normalize=True)

    @torch.no_grad()
    def preprocess_visualization(self, images: torch.Tensor) -> torch.Tensor:
        """[-1,1] autoencoder output -> de-normalised, clipped [0,1], NCHW-contiguous fp32 (base_autoencoder.py:52-61)."""
        return ops.nhwc_to_images(images, scale=0.5, shift=0.5, clamp=(0.0, 1.0))

    @abstractmethod
    def get_tokens(self, images: torch.Tensor) -> torch.Tensor: ...

    @abstractmethod
    def quantize(self, images: torch.Tensor) -> torch.Tensor: ...

    @abstractmethod
    def reconstruct(self, images: torch.Tensor) -> torch.Tensor: ...

    @abstractmethod
    def reconstruct_from_tokens(self, tokens: torch.Tensor) -> torch.Tensor: ...
