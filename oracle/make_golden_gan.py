"""Generate tests/golden/gan_*.npz by EXECUTING THE REFERENCE'S OWN loss-head modules (test infrastructure only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_gan.py

LPIPS: the pretrained VGG16 / lin weights cannot be downloaded here (no network), so torchvision's vgg16 is built with
weights=None and the lin weights are seeded torch.rand: parity is structural, with identical
random weights on both sides (the tests rebuild them from the same seed).  Discriminator: the reference class as is
(CPU -> its pure-torch `_bias_act_ref` / `_upfirdn2d_ref` paths), weights from torch.randn in constructor order."""
from __future__ import annotations

import os
import sys
from collections import OrderedDict

import numpy as np
import torch

REF = os.environ.get('VQ_REF_PATH', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')
LPIPS_CH = [64, 128, 256, 512, 512]


LPIPS_CH_ALEX = [64, 192, 384, 256, 256]


def ref_lpips(seed: int, net_type: str = 'vgg'):
    import torchvision
    import vqvae.modules.loss.lpips_pytorch.modules.networks as nets
    import vqvae.modules.loss.lpips_pytorch.modules.lpips as lp
    if not hasattr(ref_lpips, '_tv'):
        ref_lpips._tv = (torchvision.models.vgg16, torchvision.models.alexnet)
    _vgg, _alex = ref_lpips._tv
    nets.models.vgg16 = lambda weights=None, **kw: _vgg(weights=None)
    nets.models.alexnet = lambda *a, **kw: _alex(weights=None)            # the reference calls alexnet(True): no network here
    chans = LPIPS_CH if net_type == 'vgg' else LPIPS_CH_ALEX
    lp.get_state_dict = lambda net_type='alex', version='0.1': OrderedDict(
        (f'{i}.1.weight', torch.rand(1, c, 1, 1)) for i, c in enumerate(chans))
    torch.manual_seed(seed)
    return lp.LPIPS(net_type)


def lpips_case(seed=11, B=2, S=64, net_type='vgg'):
    m = ref_lpips(seed, net_type).eval()
    x = torch.rand(B, 3, S, S) * 2 - 1
    y = (torch.rand(B, 3, S, S) * 2 - 1).requires_grad_()
    out = m(x, y)
    out.backward()
    feats = m.net(y.detach())
    # the same computation in float64: ReLU / max-pool routing makes the fp32 gradient fragile at random init (the reference's
    # own fp32 result is ~0.7% away from fp64), so the tests bound the kernel's error by the reference's own fp32 error
    md = ref_lpips(seed, net_type).eval().double()
    yd = y.detach().double().requires_grad_()
    md(x.double(), yd).backward()
    return {'grad_y_f64': yd.grad.numpy(),'loss': np.float32(out.item()), 'grad_y': y.grad.numpy(), 'x': x.numpy(), 'y': y.detach().numpy(),
            'feat_sums': np.array([float(f.double().sum()) for f in feats]),
            'w0_sum': np.float64(m.net.layers[0].weight.double().sum().item())}


def disc_case(seed=21, B=4, S=64):
    from vqvae.modules.loss.stylegan2_discriminator.discriminator import Discriminator
    torch.manual_seed(seed)
    d = Discriminator(S).train()
    img = (torch.rand(B, 3, S, S) * 2 - 1).requires_grad_()
    logits = d(img)
    w = torch.linspace(-1.0, 1.0, B).reshape(B, 1)
    loss = (logits * w).sum() + torch.nn.functional.softplus(logits).mean()
    loss.backward()
    names, norms = [], []
    for n, p in d.named_parameters():
        names.append(n); norms.append(float(p.grad.double().norm()))
    # float64 evaluation of the same step: leaky-ReLU sign flips on 1-ulp forward differences make the fp32 gradients
    # fragile; the tests bound the kernel error by the reference's own fp32-vs-fp64 error
    dd = Discriminator(S).train().double()
    dd.load_state_dict({k: v.double() for k, v in d.state_dict().items()})
    for mod in dd.modules():                                   # conv2d_resample asserts float32 FIR filters
        if hasattr(mod, 'resample_filter'):
            mod.resample_filter = mod.resample_filter.float()
    imgd = img.detach().double().requires_grad_()
    ld = dd(imgd)
    ((ld * w.double()).sum() + torch.nn.functional.softplus(ld).mean()).backward()
    norms64 = [float(p.grad.norm()) for _, p in dd.named_parameters()]
    return {'grad_img_f64': imgd.grad.numpy(), 'grad_norms_f64': np.array(norms64),
            'grad_b64_conv0_w_f64': dd.b64.conv0.weight.grad.numpy()[:8], 'grad_b4_out_w_f64': dd.b4.out.weight.grad.numpy(),'logits': logits.detach().numpy(), 'loss': np.float32(loss.item()), 'grad_img': img.grad.numpy(),
            'img': img.detach().numpy(), 'grad_names': np.array(names), 'grad_norms': np.array(norms),
            'grad_b64_conv0_w': d.b64.conv0.weight.grad.numpy()[:8], 'grad_b4_out_w': d.b4.out.weight.grad.numpy(),
            'w_abs_sum': np.float64(sum(float(p.double().abs().sum()) for p in d.parameters()))}


def disc_r1_case(seed=21, B=4, S=64, cost=10.0):
    """loss.py:98-112,144-164 on the reference discriminator: non-saturating d_loss on (real, fake) + R1 penalty."""
    from vqvae.modules.loss.stylegan2_discriminator.discriminator import Discriminator
    from vqvae.modules.loss.loss import discriminator_loss
    torch.manual_seed(seed)
    d = Discriminator(S).train()
    real = (torch.rand(B, 3, S, S) * 2 - 1)
    fake = (torch.rand(B, 3, S, S) * 2 - 1)

    def run(disc, real, fake):
        real = real.clone().requires_grad_(True)
        lr = disc(real)
        lf = disc(fake)
        d_loss = discriminator_loss(lr, lf, loss_type='non-saturating')
        (g,) = torch.autograd.grad(outputs=lr.sum(), inputs=real, create_graph=True)
        r1 = cost * g.pow(2).view(g.shape[0], -1).sum(1).mean()
        (d_loss + r1).backward()
        return lr, d_loss, r1, g

    lr, d_loss, r1, g = run(d, real, fake)
    names = [n for n, _ in d.named_parameters()]
    norms = [float(p.grad.double().norm()) for _, p in d.named_parameters()]
    # R1-only gradients (d_loss excluded) isolate the second-order path
    d.zero_grad()
    real2 = real.clone().requires_grad_(True)
    (g2,) = torch.autograd.grad(outputs=d(real2).sum(), inputs=real2, create_graph=True)
    (cost * g2.pow(2).view(B, -1).sum(1).mean()).backward()
    norms_r1 = [float(p.grad.double().norm()) if p.grad is not None else 0.0 for _, p in d.named_parameters()]
    r1_b64_conv0 = d.b64.conv0.weight.grad.numpy()[:8].copy()
    r1_b4_fc_b = d.b4.fc.bias.grad.numpy().copy()
    dd = Discriminator(S).train().double()
    dd.load_state_dict({k: v.double() for k, v in d.state_dict().items()})
    for mod in dd.modules():
        if hasattr(mod, 'resample_filter'):
            mod.resample_filter = mod.resample_filter.float()
    _, d_loss64, r164, g64 = run(dd, real.double(), fake.double())
    norms64 = [float(p.grad.norm()) for _, p in dd.named_parameters()]
    return {'real': real.numpy(), 'fake': fake.numpy(), 'logits_real': lr.detach().numpy(), 'd_loss': np.float32(d_loss.item()),
            'r1': np.float32(r1.item()), 'r1_f64': np.float64(r164.item()), 'grad_real': g.detach().numpy(),
            'grad_real_f64': g64.detach().numpy(), 'grad_names': np.array(names), 'grad_norms': np.array(norms),
            'grad_norms_f64': np.array(norms64), 'grad_norms_r1_only': np.array(norms_r1),
            'r1_only_grad_b64_conv0_w': r1_b64_conv0, 'r1_only_grad_b4_fc_b': r1_b4_fc_b, 'cost': np.float32(cost)}


def main():
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    torch.set_num_threads(8)
    res = lpips_case()
    np.savez_compressed(os.path.join(OUT, 'gan_lpips_vgg.npz'), **res)
    print('lpips', res['loss'], res['feat_sums'])
    res = lpips_case(seed=13, B=2, S=96, net_type='alex')
    np.savez_compressed(os.path.join(OUT, 'gan_lpips_alex.npz'), **res)
    print('lpips alex', res['loss'], res['feat_sums'])
    res = disc_case()
    np.savez_compressed(os.path.join(OUT, 'gan_discriminator.npz'), **res)
    print('disc', res['logits'].ravel(), res['loss'])
    res = disc_r1_case()
    np.savez_compressed(os.path.join(OUT, 'gan_discriminator_r1.npz'), **res)
    print('disc r1', res['d_loss'], res['r1'], res['r1_f64'], res['grad_norms'][:4], res['grad_norms_r1_only'][:4])


if __name__ == '__main__':
    main()
