"""Import the reference's OWN `vqvae.model.VQVAE` in the build container  --  TEST INFRASTRUCTURE ONLY.

`vqvae/model.py` imports pytorch_lightning, scheduling_utils, wandb, torchmetrics and (through
abstract_modules/base_autoencoder.py) kornia; none of them is installed here and there is no network.  This module puts
minimal stand-ins for exactly the names the reference touches into `sys.modules`, so that the reference class itself --
its constructor order, `forward`, `training_step` (branch A is executed verbatim), `on_train_batch_start` and
`configure_optimizers` with the relative-name collision (defect B2) -- runs unmodified on the CPU.  Used only by
`oracle/make_golden_step.py` to produce the committed fixtures; nothing here travels to the GPU box's run-time paths.

What the stand-ins do NOT pin (stated in the fixtures' consumers as well):
  * scheduling_utils.schedulers_cpp -- un-vendored, un-pinned third party: the schedule classes are the oracle's restatement
    (semantics inferred from the call sites vqvae/model.py:175-224)  -> "parity unpinned";
  * kornia augmentations -- identity here (the parity runs bypass augmentation); Normalize / Denormalize are
    the documented (x - mean) / std and x * std + mean;
  * LPIPS pretrained weights -- not downloadable: torchvision trunks are built with weights=None and the lin weights are
    seeded torch.rand, the same stream the tests rebuild.
"""
from __future__ import annotations

import os
import sys
import types
from collections import OrderedDict

import torch
from torch import nn

REF = os.environ.get('VQ_REF_PATH', '/root/reference')
LPIPS_CH = {'vgg': [64, 128, 256, 512, 512], 'alex': [64, 192, 384, 256, 256]}


class _LightningModule(nn.Module):
    """the attributes / methods of pl.LightningModule that vqvae/model.py uses"""

    def __init__(self):
        super().__init__()
        self.trainer = None
        self.current_epoch = 0
        self.automatic_optimization = True
        self.logged = {}

    def log(self, name, value, **kwargs):
        self.logged[name] = value

    def optimizers(self):
        opts = self.trainer.optimizers
        return opts if len(opts) > 1 else opts[0]

    def manual_backward(self, loss, *args, **kwargs):
        loss.backward(*args, **kwargs)


class _Trainer:
    def __init__(self, num_training_batches: int):
        self.num_training_batches = num_training_batches
        self.optimizers = []


class _Identity(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x):
        return x


class _Normalize(nn.Module):
    def __init__(self, mean, std):
        super().__init__()
        self.mean, self.std = mean.reshape(1, -1, 1, 1), std.reshape(1, -1, 1, 1)

    def forward(self, x):
        return (x - self.mean) / self.std


class _Denormalize(_Normalize):
    def forward(self, x):
        return x * self.std + self.mean


class _Sequential(nn.Module):
    def __init__(self, *mods, **k):
        super().__init__()

    def forward(self, x):
        return x


def _scheduler_classes():
    from oracle import vqvae_oracle as orc

    class _S:
        def destroy(self):
            pass

    class LinearScheduler(_S):
        def __init__(self, a, b, va, vb):
            self.args = (a, b, va, vb)

        def step(self, i):
            return orc.linear_schedule(i, *self.args)

    class CosineScheduler(_S):
        def __init__(self, a, b, va, vb):
            self.args = (a, b, va, vb)

        def step(self, i):
            return orc.cosine_schedule(i, *self.args)

    class LinearCosineScheduler(_S):
        def __init__(self, a, b, va, vb, th):
            self.args = (a, b, va, vb, th)

        def step(self, i):
            return orc.linear_cosine_schedule(i, *self.args)

    return LinearScheduler, CosineScheduler, LinearCosineScheduler


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs() -> None:
    if 'pytorch_lightning' not in sys.modules:
        _module('pytorch_lightning', LightningModule=_LightningModule)
    lin, cos, lincos = _scheduler_classes()
    _module('scheduling_utils')
    _module('scheduling_utils.schedulers_cpp', LinearScheduler=lin, CosineScheduler=cos, LinearCosineScheduler=lincos)
    _module('wandb', Image=lambda x: x)
    _module('torchmetrics', MeanSquaredError=_Identity)
    _module('torchmetrics.image')
    _module('torchmetrics.image.fid', FrechetInceptionDistance=_Identity)
    _module('torchmetrics.image.ssim', StructuralSimilarityIndexMeasure=_Identity)
    _module('torchmetrics.image.psnr', PeakSignalNoiseRatio=_Identity)
    _module('kornia')
    _module('kornia.augmentation', AugmentationSequential=_Sequential, Denormalize=_Denormalize, Normalize=_Normalize,
            RandomHorizontalFlip=_Identity, RandomResizedCrop=_Identity)


def patch_lpips_offline() -> None:
    """seeded-random trunks and lin weights instead of the downloads"""
    import torchvision
    import vqvae.modules.loss.lpips_pytorch.modules.networks as nets
    import vqvae.modules.loss.lpips_pytorch.modules.lpips as lp
    if not hasattr(patch_lpips_offline, '_tv'):
        patch_lpips_offline._tv = (torchvision.models.vgg16, torchvision.models.alexnet)
    _vgg, _alex = patch_lpips_offline._tv
    nets.models.vgg16 = lambda weights=None, **kw: _vgg(weights=None)
    nets.models.alexnet = lambda *a, **kw: _alex(weights=None)
    lp.get_state_dict = lambda net_type='alex', version='0.1': OrderedDict(
        (f'{i}.1.weight', torch.rand(1, c, 1, 1)) for i, c in enumerate(LPIPS_CH[net_type]))


def reference_vqvae_class():
    """-> the reference's vqvae.model.VQVAE (unmodified source, executed from /root/reference)"""
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    install_stubs()
    patch_lpips_offline()
    from vqvae.model import VQVAE
    return VQVAE


def make_trainer(num_training_batches: int) -> _Trainer:
    return _Trainer(num_training_batches)
