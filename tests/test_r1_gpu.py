"""GPU parity of the R1 gradient penalty (loss.py:98-112): the gradient OF A GRADIENT through the discriminator kernels --
conv dgrad as a bilinear autograd function, leaky-ReLU masks, FIR / decimation adjoint pairs, minibatch-stddev -- against a
float64 torch evaluation of single layers and against a fixture produced by the reference discriminator
(oracle/make_golden_gan.py: disc_r1_case)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import common as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def V():
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    pkg.lib.load()
    pkg.set_precision('strict')
    return pkg


def cl(t):
    return t.cuda().contiguous(memory_format=torch.channels_last)


def r1_of(y, x, cost=10.0):
    (g,) = torch.autograd.grad(y.sum(), x, create_graph=True)
    return cost * g.pow(2).reshape(g.shape[0], -1).sum(1).mean(), g


@pytest.mark.parametrize('n,h,ci,co,k,stride', [(2, 8, 16, 24, 3, 1), (2, 9, 8, 12, 3, 2), (3, 6, 3, 32, 1, 1)])
def test_conv_lrelu_stack_second_order_exact(V, n, h, ci, co, k, stride):
    """two conv + bias + lrelu layers: penalty value, input gradient and d(penalty)/d(weights) against float64 torch."""
    torch.manual_seed(1)
    pad = k // 2 if stride == 1 else 0
    x = torch.randn(n, ci, h, h).double()
    w1 = (torch.randn(co, ci, k, k) / (ci * k * k) ** 0.5).double(); b1 = (torch.randn(co) * 0.3).double()
    w2 = (torch.randn(8, co, 3, 3) / (co * 9) ** 0.5).double(); b2 = (torch.randn(8) * 0.3).double()
    xo = x.clone().requires_grad_(); po = [t.clone().requires_grad_() for t in (w1, b1, w2, b2)]
    yo = F.leaky_relu(F.conv2d(F.leaky_relu(F.conv2d(xo, po[0], po[1], padding=pad, stride=stride), 0.2) * 1.25, po[2], po[3], padding=1), 0.2)
    ro, go = r1_of(yo, xo)
    (ro + yo.square().mean()).backward()
    xg = cl(x.float()).requires_grad_(); pg = [t.float().cuda().requires_grad_() for t in (w1, b1, w2, b2)]
    h1 = V.ops.conv2d(xg, pg[0], pg[1], None, pad=pad, stride=stride, act=V.lib.ACT_LRELU, alpha=0.2, gain=1.25)
    yg = V.ops.conv2d(h1, pg[2], pg[3], None, pad=1, act=V.lib.ACT_LRELU, alpha=0.2, gain=1.0)
    with V.ops.no_weight_gradients():
        (gg,) = torch.autograd.grad(yg.sum(), xg, create_graph=True)
    rg = 10.0 * gg.pow(2).reshape(n, -1).sum(1).mean()
    (rg + yg.square().mean()).backward()
    assert C.rel_err(gg.detach(), go.detach()) < 1e-5
    assert abs(float(rg) - float(ro)) < 1e-5 * abs(float(ro))
    for a, b in zip(pg, po):
        assert C.rel_err(a.grad, b.grad) < 2e-5
    assert C.rel_err(xg.grad, xo.grad) < 2e-5


def test_weight_gradients_refuse_to_be_silently_first_order(V):
    x = cl(torch.randn(2, 8, 6, 6)).requires_grad_(); w = torch.randn(8, 8, 3, 3, device='cuda', requires_grad=True)
    y = V.ops.conv2d(x, w, None, None, pad=1)
    with pytest.raises(V.lib.VQBError):
        torch.autograd.grad(y.sum(), x, create_graph=True)


def test_resample_and_mbstd_second_order(V):
    """FIR(pad 1, down 2) -> 1x1 conv, FIR(pad 2) -> stride-2 conv, minibatch-stddev -> conv: every adjoint pair and the
    composite second-order path of the group statistic against float64 torch."""
    from vqvae_vqgan_pytorch_lightning_b200 import ops_gan
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.discriminator import setup_filter
    torch.manual_seed(2)
    n, c, h = 4, 8, 8
    f = setup_filter().double()
    x = torch.randn(n, c, h, h).double()
    w1 = (torch.randn(c, c, 3, 3) / (c * 9) ** 0.5).double(); w2 = (torch.randn(6, c + 1, 3, 3) / (c * 9) ** 0.5).double()
    xo, w1o, w2o = x.clone().requires_grad_(), w1.clone().requires_grad_(), w2.clone().requires_grad_()
    t = F.conv2d(F.pad(xo, [2] * 4), f[None, None].repeat(c, 1, 1, 1), groups=c)
    t = F.leaky_relu(F.conv2d(t, w1o, stride=2), 0.2)                                   # [n, c, 4, 4]
    g_ = 4
    s = t.reshape(g_, -1, 1, c, 4, 4); s = s - s.mean(0); s = (s.square().mean(0) + 1e-8).sqrt().mean([2, 3, 4]).reshape(-1, 1, 1, 1).repeat(g_, 1, 4, 4)
    yo = F.leaky_relu(F.conv2d(torch.cat([t, s], 1), w2o, padding=1), 0.2)
    ro, go = r1_of(yo, xo)
    ro.backward()
    xg, w1g, w2g = cl(x.float()).requires_grad_(), w1.float().cuda().requires_grad_(), w2.float().cuda().requires_grad_()
    tg = ops_gan.fir4(xg, 2, 1)
    tg = V.ops.conv2d(tg, w1g, None, None, pad=0, stride=2, act=V.lib.ACT_LRELU, alpha=0.2, gain=1.0)
    yg = V.ops.conv2d(ops_gan.mbstd(tg, 4), w2g, None, None, pad=1, act=V.lib.ACT_LRELU, alpha=0.2, gain=1.0)
    with V.ops.no_weight_gradients():
        (gg,) = torch.autograd.grad(yg.sum(), xg, create_graph=True)
    rg = 10.0 * gg.pow(2).reshape(n, -1).sum(1).mean()
    rg.backward()
    assert C.rel_err(gg.detach(), go.detach()) < 1e-5 and abs(float(rg) - float(ro)) < 1e-5 * abs(float(ro))
    assert C.rel_err(w1g.grad, w1o.grad) < 5e-5 and C.rel_err(w2g.grad, w2o.grad) < 5e-5


def test_discriminator_r1_matches_reference_fixture(V):
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.discriminator import Discriminator
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.loss import discriminator_loss
    g = C.golden('gan_discriminator_r1')
    torch.manual_seed(21)
    d = Discriminator(64).cuda().train()
    real = cl(torch.from_numpy(g['real'])).requires_grad_(); fake = cl(torch.from_numpy(g['fake']))
    lr, lf = d(real), d(fake)
    d_loss = discriminator_loss(lr, lf, loss_type='non-saturating')
    with V.ops.no_weight_gradients():
        (gr,) = torch.autograd.grad(lr.sum(), real, create_graph=True)
    r1 = float(g['cost']) * gr.pow(2).reshape(gr.shape[0], -1).sum(1).mean()
    (d_loss + r1).backward()
    assert C.rel_err(lr, g['logits_real']) < 1e-4 and abs(float(d_loss) - float(g['d_loss'])) < 1e-4
    # leaky-ReLU sign flips (see test_discriminator_matches_reference_fixture) bound how close any fp32 evaluation gets; the
    # reference's own fp32 result is compared with its float64 evaluation for scale
    # (measured: 5.6e-3 here, i.e. a handful of flipped slopes; the single-layer tests above hold the kernels to 1e-5)
    assert C.rel_err(gr.detach(), g['grad_real_f64']) < 1.5e-2
    assert abs(float(r1) - float(g['r1_f64'])) < 5e-3 * float(g['r1_f64'])
    ref_norm = dict(zip(g['grad_names'].tolist(), g['grad_norms_f64'].tolist()))
    for n, p in d.named_parameters():
        assert abs(float(p.grad.double().norm()) - ref_norm[n]) <= 1e-2 * ref_norm[n] + 1e-7, n


def test_r1_only_gradients_match_reference_fixture(V):
    """d(R1)/d(theta) alone (no d_loss): isolates the second-order path, including the bias gradients that exist only through
    the minibatch-stddev statistic."""
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.discriminator import Discriminator
    g = C.golden('gan_discriminator_r1')
    torch.manual_seed(21)
    d = Discriminator(64).cuda().train()
    real = cl(torch.from_numpy(g['real'])).requires_grad_()
    with V.ops.no_weight_gradients():
        (gr,) = torch.autograd.grad(d(real).sum(), real, create_graph=True)
    (float(g['cost']) * gr.pow(2).reshape(gr.shape[0], -1).sum(1).mean()).backward()
    ref_norm = dict(zip(g['grad_names'].tolist(), g['grad_norms_r1_only'].tolist()))
    for n, p in d.named_parameters():
        got = float(p.grad.double().norm()) if p.grad is not None else 0.0
        tol = 2e-2 if n.endswith('bias') else 1e-2
        assert abs(got - ref_norm[n]) <= tol * ref_norm[n] + 1e-8, (n, got, ref_norm[n])
    assert C.rel_err(d.b64.conv0.weight.grad[:8], g['r1_only_grad_b64_conv0_w']) < 2e-2
    # biases after the minibatch-stddev layer only enter R1 through the (locally constant) lrelu masks: exactly zero
    assert float(np.abs(g['r1_only_grad_b4_fc_b']).max()) == 0.0
    assert d.b4.fc.bias.grad is None or float(d.b4.fc.bias.grad.abs().max()) == 0.0


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_vqgan_step_with_r1(V, precision):
    """the example VQGAN configuration's adversarial block (r1_reg_weight 10, every 16 steps -- gumbel_vqgan.yaml:35-36)."""
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    if precision == 'fast' and not V.lib.load().vqb_device_supports_tcgen05():
        pytest.skip('needs sm_100')
    V.set_precision(precision)
    try:
        torch.manual_seed(3)
        adv = dict(start_epoch=0, loss_type='non-saturating', g_weight=0.1, use_adaptive=False, r1_reg_weight=10., r1_reg_every=2)
        model = V.VQVAE(64, dict(channels=128, num_res_blocks=1, channel_multipliers=[1, 2]),
                        dict(num_embeddings=64, embedding_dim=64, type='ema', params=dict(commitment_cost=0.25, decay=0.95, epsilon=1e-5),
                             reinit_every_n_epochs=None),
                        dict(l1_weight=0.8, l2_weight=0.2, perc_weight=1.0, adversarial_params=adv),
                        dict(lr=1e-4, betas=[0.0, 0.99], eps=1e-8, weight_decay=1e-4, warmup_epochs=None, decay_epochs=None),
                        pretrained_lpips=False).cuda().train()
        tr = Trainer(max_epochs=1, num_training_batches=4)
        tr.attach(model); model.on_train_start()
        x = torch.rand(4, 3, 64, 64, device='cuda')
        seen = []
        for i in range(3):
            loss = tr.run_step(x, i)
            seen.append(float(torch.as_tensor(model.logged['r1_penalty'])))
        assert torch.isfinite(loss).all() and np.isfinite(seen).all()
        assert seen[0] > 0 and seen[1] == 0 and seen[2] > 0            # steps 0 and 2 carry the penalty
    finally:
        V.set_precision('strict')
