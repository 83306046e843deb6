"""Pre/post-processing base class -- surface of the reference's BaseVQVAE
(vqvae/modules/abstract_modules/base_autoencoder.py:6-93)."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch

from ... import ops


class BaseVQVAE(ABC):

    def __init__(self, image_size: int):
        self.image_size = image_size
        # the reference applies kornia RandomResizedCrop(0.7-1, ratio 1) + RandomHorizontalFlip when training
        # (base_autoencoder.py:17-22,45-46).  kornia is un-pinned and absent here (PARITY UNPINNED); augmentation is
        # the "next" row 8f-3 of the scope table and is off unless a callable is installed here.
        self.training_augmentations = None
        self.scheduler = None
        self.train_epoch_usage_count = None
        self.val_epoch_usage_count = None

    @torch.no_grad()
    def preprocess_batch(self, images: torch.Tensor, training: bool = False) -> torch.Tensor:
        """images [B,C,H,W] fp32 in [0,1] -> clamp, (optional augmentation), (x-0.5)/0.5; returned channels-last fp32.
        One fused kernel (vqb_nchw_to_nhwc) instead of clamp + Normalize (base_autoencoder.py:41-50)."""
        if training and self.training_augmentations is not None:
            images = self.training_augmentations(images)
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
