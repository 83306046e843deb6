"""VQGAN loss heads (reference: vqvae/modules/loss/loss.py): L1/L2/LPIPS reconstruction terms, hinge / non-saturating
adversarial losses, adaptive generator weight.  The image-sized reductions run in libvqgan_b200 kernels; the arithmetic on
the [B,1] logits and on scalars is plain tensor glue."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as functional
from torch.autograd import grad

from ... import ops, ops_gan
from .discriminator import Discriminator
from .lpips import LPIPS


def generator_loss(logits: torch.Tensor, loss_type: str = 'hinge'):
    """loss.py:11-26"""
    if loss_type == 'hinge':
        return -torch.mean(logits)
    if loss_type == 'non-saturating':
        return functional.binary_cross_entropy_with_logits(logits, target=torch.ones_like(logits))
    raise ValueError(f'unknown loss_type: {loss_type}')


def discriminator_loss(logits_real: torch.Tensor, logits_fake: torch.Tensor, loss_type: str = 'hinge'):
    """loss.py:29-51"""
    if loss_type == 'hinge':
        real_loss = functional.relu(1.0 - logits_real)
        fake_loss = functional.relu(1.0 + logits_fake)
    elif loss_type == 'non-saturating':
        real_loss = functional.binary_cross_entropy_with_logits(logits_real, target=torch.ones_like(logits_real), reduction='none')
        fake_loss = functional.binary_cross_entropy_with_logits(logits_fake, target=torch.zeros_like(logits_fake), reduction='none')
    else:
        raise ValueError(f'unknown loss_type: {loss_type}')
    return torch.mean(real_loss + fake_loss)


class VQLPIPSWithDiscriminator(nn.Module):
    """loss.py:54-164"""

    def __init__(self, image_size: int, l1_weight: float, l2_weight: float, perc_weight: float, adversarial_conf: dict,
                 pretrained_lpips: bool = True):
        super().__init__()
        self.l1_weight = l1_weight
        self.l2_weight = l2_weight
        self.perceptual_loss = LPIPS(net_type='vgg', pretrained=pretrained_lpips)
        self.perceptual_weight = perc_weight
        self.discriminator = Discriminator(image_size)
        self.adversarial_start_epoch = adversarial_conf['start_epoch']
        self.adversarial_loss_type = adversarial_conf['loss_type']
        self.generator_weight = adversarial_conf['g_weight']
        self.use_adaptive_g_weight = adversarial_conf['use_adaptive']
        self.r1_regularization_cost = adversarial_conf['r1_reg_weight']
        self.r1_regularization_every = adversarial_conf['r1_reg_every']

    def calculate_adaptive_weight(self, nll_loss, g_loss, last_layer):
        """loss.py:80-96 (the caller passes the perceptual loss as `nll_loss`, :131)"""
        with ops.no_grad_sink():           # these parameter gradients are RETURNED, not accumulated into last_layer.grad
            nll_grads = grad(nll_loss, last_layer, grad_outputs=torch.ones_like(nll_loss), retain_graph=True)[0].detach()
            g_grads = grad(g_loss, last_layer, grad_outputs=torch.ones_like(g_loss), retain_graph=True)[0].detach()
        adaptive_weight = torch.norm(nll_grads, p=2) / (torch.norm(g_grads, p=2) + 1e-8)
        adaptive_weight = torch.clamp(adaptive_weight, 0.0, 1e4).detach()
        return adaptive_weight * self.generator_weight

    def calculate_r1_regularization_term(self, logits_real, images, compute_r1: bool):
        """loss.py:98-112: cost * mean_b sum (d sum(logits_real) / d images)^2, with the graph of that backward pass recorded
        so that the penalty itself is differentiable in the discriminator weights.  Every discriminator op is twice
        differentiable on the kernels (ops.ConvDgradFn / ActBwdFn and the adjoint pairs in ops_gan)."""
        if compute_r1:
            with ops.no_weight_gradients():
                (gradients,) = torch.autograd.grad(outputs=logits_real.sum(), inputs=images, create_graph=True)
            return self.r1_regularization_cost * gradients.pow(2).reshape(gradients.shape[0], -1).sum(1).mean()
        return 0.

    def forward_autoencoder(self, quantizer_loss, images, reconstructions, current_epoch: int, last_layer):
        l2_loss, l1_loss = ops.mse_l1(reconstructions, images)
        p_loss = self.perceptual_loss(images, reconstructions)
        nll_loss = l1_loss * self.l1_weight + l2_loss * self.l2_weight + p_loss * self.perceptual_weight
        if current_epoch >= self.adversarial_start_epoch:
            # the reference leaves D trainable here and throws its gradients away (defect B11); freezing D for this pass
            # yields identical autoencoder gradients without the wasted weight-gradient kernels
            d_params = [p for p in self.discriminator.parameters() if p.requires_grad]
            for p in d_params:
                p.requires_grad_(False)
            try:
                logits_fake = self.discriminator(reconstructions)
            finally:
                for p in d_params:
                    p.requires_grad_(True)
            g_loss = generator_loss(logits_fake, loss_type=self.adversarial_loss_type)
            if self.training and self.use_adaptive_g_weight:
                g_weight = self.calculate_adaptive_weight(p_loss, g_loss, last_layer=last_layer)
            else:
                g_weight = self.generator_weight
            loss = nll_loss + g_loss * g_weight + quantizer_loss
        else:
            g_loss = torch.zeros_like(nll_loss, requires_grad=False)
            g_weight = 0.
            loss = nll_loss + quantizer_loss
        return loss, l1_loss, l2_loss, p_loss, g_loss, g_weight

    def forward_discriminator(self, images, reconstructions, current_epoch: int, current_step: int):
        if current_epoch >= self.adversarial_start_epoch:
            compute_r1 = (self.training and current_step % self.r1_regularization_every == 0 and
                          self.r1_regularization_cost is not None)
            if compute_r1:
                images = images.detach().requires_grad_(True)
            if compute_r1:
                with ops_gan.second_order():            # this pass is differentiated twice (R1): twice-differentiable layer routes
                    logits_real = self.discriminator(images)
            else:
                logits_real = self.discriminator(images)
            logits_fake = self.discriminator(reconstructions.detach())
            d_loss = discriminator_loss(logits_real, logits_fake, loss_type=self.adversarial_loss_type)
            r1_term = self.calculate_r1_regularization_term(logits_real, images, compute_r1)
            loss = d_loss + r1_term
        else:
            d_loss = torch.zeros((1,), device=images.device)
            r1_term = 0.
            loss = None
        return loss, d_loss, r1_term


class VQLPIPS(nn.Module):
    """loss.py:167-199 (ablation: LPIPS without discriminator), AlexNet trunk as in the reference (`net_type='alex'`)."""

    def __init__(self, l1_weight: float, l2_weight: float, perc_weight: float, net_type: str = 'alex', pretrained_lpips: bool = True):
        super().__init__()
        self.l1_weight = l1_weight
        self.l2_weight = l2_weight
        self.perceptual_loss = LPIPS(net_type=net_type, pretrained=pretrained_lpips)
        self.perceptual_weight = perc_weight

    def forward(self, quantizer_loss, images, reconstructions):
        l2_loss, l1_loss = ops.mse_l1(reconstructions, images)
        p_loss = self.perceptual_loss(images, reconstructions)
        nll_loss = l1_loss * self.l1_weight + l2_loss * self.l2_weight + p_loss * self.perceptual_weight
        return quantizer_loss + nll_loss, l1_loss, l2_loss, p_loss
