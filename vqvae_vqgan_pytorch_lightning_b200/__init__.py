"""vqgan-b200: Blackwell-native VQ-VAE / VQGAN training step.

Import never builds anything: `__graft_entry__.build()` (or `python -m vqvae_vqgan_pytorch_lightning_b200.build`)
compiles libvqgan_b200.so; every op raises lib.VQBError when the library is missing (no CPU / eager fallback).
"""
from . import lib, ops  # noqa: F401
from .model import VQVAE  # noqa: F401
from .ops import set_precision, get_precision  # noqa: F401

__all__ = ['VQVAE', 'lib', 'ops', 'set_precision', 'get_precision']
