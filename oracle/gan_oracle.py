"""CPU oracle for the VQGAN loss heads and the full optimisation step  --  TEST INFRASTRUCTURE ONLY.

Functional (state-dict driven) fp32 restatement of
  * LPIPS (vqvae/modules/loss/lpips_pytorch/modules/{lpips,networks,utils}.py),
  * the StyleGAN2 discriminator (vqvae/modules/loss/stylegan2_discriminator/discriminator.py and the pure-torch reference
    paths of its ops: utils/ops/{conv2d_resample,upfirdn2d,bias_act}.py),
  * the loss heads (vqvae/modules/loss/loss.py:11-199),
  * one whole optimisation step of every branch of VQVAE.training_step incl. the schedules and AdamW with the reference's
    parameter grouping (vqvae/model.py:202-295, 372-440).
Only tests/, __graft_entry__.smoke() and bench.py's baseline legs (cpu_baseline, --impl reference, torch_eager_gpu_baseline) may
import this file.

Pinning: tests/golden/step_*.npz are produced by oracle/make_golden_step.py, which executes the reference's OWN VQVAE class
(oracle/ref_harness.py); tests/test_oracle_step.py checks this file against them on the CPU.  The schedule classes come
from the un-vendored `scheduling_utils` package: PARITY UNPINNED (restated from the call sites).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import vqvae_oracle as orc

Tensor = torch.Tensor
SD = Dict[str, Tensor]

VGG_CFG = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512)
VGG_TAPS = (4, 9, 16, 23, 30)
ALEX_TAPS = (2, 5, 8, 10, 12)


# --------------------------------------------------------------------------------------
# LPIPS
# --------------------------------------------------------------------------------------
def _unit_normalise(x: Tensor, eps: float = 1e-10) -> Tensor:
    """lpips_pytorch/modules/utils.py:6-8"""
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


def lpips_features(sd: SD, p: str, x: Tensor, net_type: str) -> List[Tensor]:
    """networks.py:51-64 (z-score, trunk, unit-normalised taps); trunks :78-97"""
    x = (x - sd[p + 'net.mean']) / sd[p + 'net.std']
    feats, i = [], 0                     # i = 1-based index of the torchvision `features` layer just applied

    def conv(x, idx, **kw):
        return F.relu(F.conv2d(x, sd[f'{p}net.layers.{idx}.weight'], sd[f'{p}net.layers.{idx}.bias'], **kw))

    if net_type == 'vgg':
        for v in VGG_CFG:
            if v == 'M':
                x = F.max_pool2d(x, 2, 2); i += 1
            else:
                x = conv(x, i, padding=1); i += 2
            if i in VGG_TAPS:
                feats.append(_unit_normalise(x))
    elif net_type == 'alex':
        x = conv(x, 0, stride=4, padding=2); feats.append(_unit_normalise(x))            # layers 1-2
        x = F.max_pool2d(x, 3, 2)
        x = conv(x, 3, padding=2); feats.append(_unit_normalise(x))                      # 4-5
        x = F.max_pool2d(x, 3, 2)
        x = conv(x, 6, padding=1); feats.append(_unit_normalise(x))                      # 7-8
        x = conv(x, 8, padding=1); feats.append(_unit_normalise(x))                      # 9-10
        x = conv(x, 10, padding=1); feats.append(_unit_normalise(x))                     # 11-12
    else:
        raise NotImplementedError(net_type)
    return feats


def lpips(sd: SD, p: str, x: Tensor, y: Tensor, net_type: str) -> Tensor:
    """lpips.py:31-38: per tap (fx-fy)^2 -> 1x1 lin (no bias) -> spatial mean; sum over taps; batch mean."""
    fx, fy = lpips_features(sd, p, x, net_type), lpips_features(sd, p, y, net_type)
    res = [F.conv2d((a - b) ** 2, sd[f'{p}lin.{i}.1.weight']).mean((2, 3), True) for i, (a, b) in enumerate(zip(fx, fy))]
    return torch.mean(torch.sum(torch.cat(res, 1), 1))


# --------------------------------------------------------------------------------------
# StyleGAN2 discriminator
# --------------------------------------------------------------------------------------
def _fir(x: Tensor, f: Tensor, pad: int) -> Tensor:
    """upfirdn2d.py:162-208 (reference path) with up=down=1: zero-pad, depthwise correlation with the (symmetric) 4x4 filter"""
    c = x.shape[1]
    return F.conv2d(F.pad(x, [pad, pad, pad, pad]), f[None, None].repeat(c, 1, 1, 1), groups=c)


def d_conv(sd: SD, name: str, x: Tensor, k: int, down: int = 1, act: str = 'lrelu', gain: float = 1.0) -> Tensor:
    """discriminator.py:164-174 + conv2d_resample.py:100-122,145-147 + bias_act.py:55-97 (lrelu 0.2, def_gain sqrt 2)."""
    w = sd[name + '.weight']
    w = w * (1.0 / math.sqrt(w.shape[1] * k * k))
    f = sd[name + '.resample_filter']
    if down == 1:
        x = F.conv2d(x, w, padding=k // 2)
    elif k == 1:
        x = F.conv2d(_fir(x, f, 1)[:, :, ::2, ::2], w)            # FIR + decimate, then 1x1
    else:
        x = F.conv2d(_fir(x, f, 2), w, stride=2)                  # pad (k//2 + 1) each side, FIR, stride-2 conv without padding
    b = sd.get(name + '.bias')
    if b is not None:
        x = x + b.reshape(1, -1, 1, 1)
    if act == 'lrelu':
        x = F.leaky_relu(x, 0.2) * (math.sqrt(2.0) * gain)
    elif gain != 1.0:
        x = x * gain
    return x


def minibatch_std(x: Tensor, group: int = 4) -> Tensor:
    """discriminator.py:277-293 (num_channels = 1)"""
    n, c, h, w = x.shape
    g = min(group, n)
    y = x.reshape(g, -1, 1, c, h, w)
    y = y - y.mean(dim=0)
    y = (y.square().mean(dim=0) + 1e-8).sqrt().mean(dim=[2, 3, 4])
    return torch.cat([x, y.reshape(-1, 1, 1, 1).repeat(g, 1, h, w)], dim=1)


def discriminator(sd: SD, p: str, img: Tensor) -> Tensor:
    """discriminator.py:237-265,328-354,404-413, architecture 'resnet', fp32 -> logits [B,1]"""
    res = img.shape[-1]
    x = None
    while res > 4:
        b = f'{p}b{res}.'
        if x is None:
            x = d_conv(sd, b + 'fromrgb', img, 1)
        y = d_conv(sd, b + 'skip', x, 1, down=2, act='linear', gain=math.sqrt(0.5))
        x = d_conv(sd, b + 'conv0', x, 3)
        x = d_conv(sd, b + 'conv1', x, 3, down=2, gain=math.sqrt(0.5))
        x = y + x
        res //= 2
    x = d_conv(sd, p + 'b4.conv', minibatch_std(x), 3)
    x = x.flatten(1)
    w = sd[p + 'b4.fc.weight']
    x = F.leaky_relu(x.matmul((w * (1.0 / math.sqrt(w.shape[1]))).t()) + sd[p + 'b4.fc.bias'], 0.2) * math.sqrt(2.0)
    w = sd[p + 'b4.out.weight']
    return torch.addmm(sd[p + 'b4.out.bias'].unsqueeze(0), x, (w * (1.0 / math.sqrt(w.shape[1]))).t())


# --------------------------------------------------------------------------------------
# loss heads (vqvae/modules/loss/loss.py)
# --------------------------------------------------------------------------------------
def generator_loss(logits: Tensor, loss_type: str) -> Tensor:
    """loss.py:11-26"""
    if loss_type == 'hinge':
        return -torch.mean(logits)
    if loss_type == 'non-saturating':
        return F.binary_cross_entropy_with_logits(logits, torch.ones_like(logits))
    raise ValueError(f'unknown loss_type: {loss_type}')


def discriminator_loss(lr_: Tensor, lf: Tensor, loss_type: str) -> Tensor:
    """loss.py:29-51"""
    if loss_type == 'hinge':
        return torch.mean(F.relu(1.0 - lr_) + F.relu(1.0 + lf))
    if loss_type == 'non-saturating':
        return torch.mean(F.binary_cross_entropy_with_logits(lr_, torch.ones_like(lr_), reduction='none') +
                          F.binary_cross_entropy_with_logits(lf, torch.zeros_like(lf), reduction='none'))
    raise ValueError(f'unknown loss_type: {loss_type}')


def forward_autoencoder(sd: SD, l_conf: dict, q_loss: Tensor, images: Tensor, recon: Tensor, epoch: int, training: bool = True):
    """loss.py:114-142 (VQLPIPSWithDiscriminator) and :185-199 (VQLPIPS when adversarial_params is None).
    -> (loss, l1, l2, p_loss, g_loss, g_weight).  The adaptive weight (:80-96) uses the PERCEPTUAL loss as `nll_loss` (:131)
    and the gradients w.r.t. decoder.conv_out.weight."""
    adv = l_conf.get('adversarial_params')
    net = 'vgg' if adv is not None else 'alex'
    l1 = (images - recon).abs().mean()
    l2 = (images - recon).pow(2).mean()
    p_loss = lpips(sd, 'criterion.perceptual_loss.', images, recon, net)
    nll = l1 * l_conf['l1_weight'] + l2 * l_conf['l2_weight'] + p_loss * l_conf['perc_weight']
    if adv is None:
        return q_loss + nll, l1, l2, p_loss, torch.zeros(1), 0.
    if epoch >= adv['start_epoch']:
        g_loss = generator_loss(discriminator(sd, 'criterion.discriminator.', recon), adv['loss_type'])
        if training and adv['use_adaptive']:
            last = sd['decoder.conv_out.weight']
            gn = torch.autograd.grad(p_loss, last, retain_graph=True)[0].detach()
            gg = torch.autograd.grad(g_loss, last, retain_graph=True)[0].detach()
            g_weight = torch.clamp(torch.norm(gn, p=2) / (torch.norm(gg, p=2) + 1e-8), 0.0, 1e4).detach() * adv['g_weight']
        else:
            g_weight = adv['g_weight']
        return nll + g_loss * g_weight + q_loss, l1, l2, p_loss, g_loss, g_weight
    return nll + q_loss, l1, l2, p_loss, torch.zeros_like(nll), 0.


def forward_discriminator(sd: SD, l_conf: dict, images: Tensor, recon: Tensor, epoch: int, step: int, training: bool = True):
    """loss.py:144-164 with the R1 term of :98-112 -> (loss | None, d_loss, r1)"""
    adv = l_conf['adversarial_params']
    if epoch < adv['start_epoch']:
        return None, torch.zeros(1), 0.
    compute_r1 = training and step % adv['r1_reg_every'] == 0 and adv['r1_reg_weight'] is not None
    images = images.detach().requires_grad_(compute_r1)
    lr_ = discriminator(sd, 'criterion.discriminator.', images)
    lf = discriminator(sd, 'criterion.discriminator.', recon.detach())
    d_loss = discriminator_loss(lr_, lf, adv['loss_type'])
    r1 = 0.
    if compute_r1:
        (g,) = torch.autograd.grad(outputs=lr_.sum(), inputs=images, create_graph=True)
        r1 = adv['r1_reg_weight'] * g.pow(2).reshape(g.shape[0], -1).sum(1).mean()
    return d_loss + r1, d_loss, r1


# --------------------------------------------------------------------------------------
# optimizers and the training loop body (vqvae/model.py)
# --------------------------------------------------------------------------------------
class AdamW:
    """torch.optim.AdamW over named tensors of `sd` in groups [(names, weight_decay)]; tensors without a gradient are skipped."""

    def __init__(self, sd: SD, groups: Sequence, lr: float, betas: Sequence[float], eps: float):
        self.sd, self.groups, self.lr, self.betas, self.eps = sd, list(groups), lr, tuple(betas), eps
        self.state = {n: [torch.zeros_like(sd[n]), torch.zeros_like(sd[n]), 0] for names, _ in self.groups for n in names}

    def zero_grad(self):
        for names, _ in self.groups:
            for n in names:
                self.sd[n].grad = None

    @torch.no_grad()
    def step(self):
        for names, wd in self.groups:
            for n in names:
                p = self.sd[n]
                if p.grad is None:
                    continue
                st = self.state[n]
                st[2] += 1
                orc.adamw_step(p, p.grad, st[0], st[1], st[2], self.lr, self.betas[0], self.betas[1], self.eps, wd)


def configure_optimizers(sd: SD, t_conf: dict, gan: bool, fix_param_groups: bool = False):
    """model.py:372-440: AE AdamW with the (decay | no-decay) split and the relative-name collision (defect B2: the
    decoder tensor of a colliding relative name replaces the encoder tensor, which is then never optimised); the
    discriminator AdamW decays every parameter (:431-433)."""
    rel = lambda pre: [n[len(pre):] for n in sd if n.startswith(pre) and sd[n].requires_grad]
    decay, no_decay = orc.adamw_groups(rel('encoder.'), rel('decoder.'), rel('quantizer.'), not fix_param_groups)
    lr, betas, eps, wd = float(t_conf['lr']), [float(b) for b in t_conf['betas']], float(t_conf['eps']), float(t_conf['weight_decay'])
    opts = [AdamW(sd, [(decay, wd), (no_decay, 0.0)], lr, betas, eps)]
    if gan:
        d_names = [n for n in sd if n.startswith('criterion.discriminator.') and sd[n].requires_grad]
        opts.append(AdamW(sd, [(d_names, wd)], lr, betas, eps))
    return opts


def step_schedules(t_conf: dict, q_cfg: dict, nb: int, step: int):
    """model.py:163-230 -> (lr, gumbel temperature | None, gumbel kl_cost | None) at `step` (PARITY UNPINNED schedules)."""
    lr = float(t_conf['lr'])
    wu, dc = t_conf.get('warmup_epochs'), t_conf.get('decay_epochs')
    if wu is not None and dc is not None:
        lr_s = orc.linear_cosine_schedule(step, 0, dc * nb, lr, lr / 2., wu * nb)
    elif wu is not None:
        lr_s = orc.linear_schedule(step, 0, wu * nb, 1e-20, lr)
    elif dc is not None:
        lr_s = orc.cosine_schedule(step, 0, dc * nb, lr, lr / 2.)
    else:
        lr_s = lr
    temp = kl = None
    if q_cfg['type'] == 'gumbel':
        temp, kl = q_cfg['temp'], q_cfg['kl_cost']
        if q_cfg.get('kl_warmup_epochs') is not None:
            kl = orc.cosine_schedule(step, 0, int(q_cfg['kl_warmup_epochs'] * nb), 0.0, q_cfg['kl_cost'])
        if q_cfg.get('temp_decay_epochs') is not None and q_cfg.get('temp_final') is not None:
            temp = orc.cosine_schedule(step, 0, int(q_cfg['temp_decay_epochs'] * nb), q_cfg['temp'], q_cfg['temp_final'])
    return lr_s, temp, kl


def train_step(sd: SD, opts: Sequence[AdamW], images01: Tensor, cfg: dict, l_conf: Optional[dict], t_conf: dict,
               epoch: int, batch_index: int, nb: int, exp_noise: Optional[Tensor] = None) -> dict:
    """One iteration of the fit loop: on_train_batch_start (model.py:202-230) + training_step (:232-295) + the optimizer
    steps (Lightning automatic optimisation for branches B / C, manual for branch A :244-264).  Mutates `sd`."""
    step = epoch * nb + batch_index
    lr, temp, kl = step_schedules(t_conf, cfg['quantizer'], nb, step)
    for o in opts:
        o.lr = lr
    cfg = dict(cfg)
    if temp is not None:
        cfg['quantizer'] = dict(cfg['quantizer'], temp=temp, kl_cost=kl)
    images = orc.normalize_images(images01)
    out = orc.forward_vqvae(sd, images, cfg, training=True, exp_noise=exp_noise)
    recon, q_loss = out['recon'], out['q_loss']
    log = {'lr': lr, 'idx': out['idx']}
    gan = l_conf is not None and l_conf.get('adversarial_params') is not None
    if l_conf is None:
        l2 = F.mse_loss(recon, images)
        ae_loss = q_loss + l2
        log.update(l1=0., l2=l2, p=0., g=0., g_weight=0., d=0., r1=0.)
    else:
        ae_loss, l1, l2, p_loss, g_loss, g_weight = forward_autoencoder(sd, l_conf, q_loss, images, recon, epoch)
        log.update(l1=l1, l2=l2, p=p_loss, g=g_loss, g_weight=g_weight, d=0., r1=0.)
    opts[0].zero_grad()
    if gan:                                                    # the generator pass leaves gradients on D too (defect B11);
        for n in sd:                                           # they are discarded by disc_opt.zero_grad() below
            if n.startswith('criterion.discriminator.'):
                sd[n].grad = None
    ae_loss.backward()
    if 'new_codebook' in out:                                  # EMA state is replaced inside the forward (vector_quantizers.py:158-169)
        with torch.no_grad():
            sd['quantizer.codebook.weight'].copy_(out['new_codebook'])
            sd['quantizer.ema_count'].copy_(out['new_ema_count'])
            sd['quantizer.ema_weight'].copy_(out['new_ema_weight'])
    opts[0].step()
    if gan:
        d_total, d_loss, r1 = forward_discriminator(sd, l_conf, images, recon, epoch, step)
        if d_total is not None:
            opts[1].zero_grad()
            d_total.backward()
            opts[1].step()
        log.update(d=d_loss, r1=r1)
    log.update(loss=ae_loss, q=q_loss)
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in log.items()}
