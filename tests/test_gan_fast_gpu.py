"""GPU: the discriminator's down-sampling layers in the bf16 fast mode -- vectorised FIR kernels, and the stride-2 3x3
convolution executed as a full-resolution stride-1 tcgen05 convolution + decimation -- against fp32 torch on bf16-rounded
operands, plus one full VQGAN step in fast mode."""
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
    if not pkg.lib.load().vqb_device_supports_tcgen05():
        pytest.skip('needs sm_100')
    pkg.set_precision('fast')
    yield pkg
    pkg.set_precision('strict')


def cl(t):
    return t.cuda().contiguous(memory_format=torch.channels_last)


def r16(t):
    return t.bfloat16().float()


def test_down_conv_layer_tensor_core_route(V):
    """Conv2dLayer(down=2, k=3): FIR(pad 2) -> stride-2 conv (no pad) -> bias -> lrelu * gain, forward and backward."""
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.discriminator import Conv2dLayer, setup_filter
    torch.manual_seed(4)
    n, ci, co, h = 2, 128, 256, 32
    layer = Conv2dLayer(ci, co, kernel_size=3, activation='lrelu', down=2)
    with torch.no_grad():
        layer.weight.copy_(r16(layer.weight)); layer.bias.copy_(torch.randn(co) * 0.1)
    x = r16(torch.randn(n, ci, h, h)); go = r16(torch.randn(n, co, h // 2, h // 2))
    f = setup_filter()
    xo = x.clone().requires_grad_(); wo = layer.weight.detach().clone().requires_grad_(); bo = layer.bias.detach().clone().requires_grad_()
    xb = F.conv2d(F.pad(xo, [2, 2, 2, 2]), f[None, None].repeat(ci, 1, 1, 1), groups=ci)
    # the fast route stores the FIR output and the scaled weights in bf16; round the reference's the same way
    # (straight-through).  Otherwise the ~0.06 % of pre-activations that change sign under that rounding flip the lrelu
    # slope (0.2 <-> 1) and alone put ~2.5 % of relative L2 error on every gradient (tools/debug_downconv.py).
    xb = xb + (r16(xb) - xb).detach()
    wq = wo * layer.weight_gain
    wq = wq + (r16(wq) - wq).detach()                              # the packed tensor-core operand is bf16(w * gain)
    yo = F.leaky_relu(F.conv2d(xb, wq, bo, stride=2), 0.2) * (np.sqrt(2) * np.sqrt(0.5))
    yo.backward(go)
    layer = layer.cuda()
    xg = cl(x).bfloat16().requires_grad_()
    yg = layer(xg, gain=np.sqrt(0.5))
    assert yg.shape == yo.shape and yg.dtype == torch.bfloat16
    yg.backward(cl(go).bfloat16())
    assert C.rel_err(yg.float(), yo) < 5e-3
    assert C.rel_err(xg.grad.float(), xo.grad) < 1e-2
    assert C.rel_err(layer.weight.grad, wo.grad) < 1e-2 and C.rel_err(layer.bias.grad, bo.grad) < 1e-2


def test_vectorised_fir_matches_reference(V):
    from vqvae_vqgan_pytorch_lightning_b200 import ops_gan
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.discriminator import setup_filter
    torch.manual_seed(5)
    f = setup_filter()
    for (pad, down, h, w, c) in ((2, 1, 16, 16, 64), (1, 2, 16, 12, 128)):
        x = r16(torch.randn(2, c, h, w)); xo = x.clone().requires_grad_()
        y = F.conv2d(F.pad(xo, [pad] * 4), f[None, None].repeat(c, 1, 1, 1), groups=c)[:, :, ::down, ::down]
        go = r16(torch.randn_like(y)); y.backward(go)
        xg = cl(x).bfloat16().requires_grad_()
        yg = ops_gan.fir4(xg, pad, down); yg.backward(cl(go).bfloat16())
        assert C.rel_err(yg.float(), y) < 4e-3 and C.rel_err(xg.grad.float(), xo.grad) < 4e-3


def test_vqgan_step_fast_mode(V):
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    torch.manual_seed(3)
    adv = dict(start_epoch=0, loss_type='hinge', g_weight=0.1, use_adaptive=False, r1_reg_weight=None, r1_reg_every=16)
    model = V.VQVAE(64, dict(channels=128, num_res_blocks=1, channel_multipliers=[1, 2]),
                    dict(num_embeddings=256, embedding_dim=64, type='gumbel',
                         params=dict(straight_through=False, temp=1.0, kl_cost=0.0086, kl_warmup_epochs=None, temp_decay_epochs=None,
                                     temp_final=None), reinit_every_n_epochs=None),
                    dict(l1_weight=0.8, l2_weight=0.2, perc_weight=1.0, adversarial_params=adv),
                    dict(lr=1e-4, betas=[0.0, 0.99], eps=1e-8, weight_decay=1e-4, warmup_epochs=None, decay_epochs=None),
                    pretrained_lpips=False).cuda().train()
    tr = Trainer(max_epochs=1, num_training_batches=4)
    tr.attach(model); model.on_train_start()
    x = torch.rand(4, 3, 64, 64, device='cuda')
    for i in range(2):
        loss = tr.run_step(x, i)
    assert torch.isfinite(loss).all()
    for k in ('train/perc_loss', 'train/gen_loss', 'train/disc_loss', 'train/quant_loss'):
        assert torch.isfinite(torch.as_tensor(model.logged[k])).all(), k


def test_lpips_only_step_fast_mode_alexnet(V):
    """branch B of training_step (VQLPIPS, AlexNet trunk as in loss.py:182) in the bf16 fast mode."""
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    torch.manual_seed(4)
    model = V.VQVAE(96, dict(channels=128, num_res_blocks=1, channel_multipliers=[1, 2]),
                    dict(num_embeddings=128, embedding_dim=64, type='ema', params=dict(commitment_cost=0.25, decay=0.95, epsilon=1e-5),
                         reinit_every_n_epochs=None),
                    dict(l1_weight=0.8, l2_weight=0.2, perc_weight=1.0, adversarial_params=None),
                    dict(lr=1e-4, betas=[0.9, 0.99], eps=1e-8, weight_decay=1e-4, warmup_epochs=None, decay_epochs=None),
                    pretrained_lpips=False).cuda().train()
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.lpips import AlexNet
    assert isinstance(model.criterion.perceptual_loss.net, AlexNet)
    tr = Trainer(max_epochs=1, num_training_batches=4)
    tr.attach(model); model.on_train_start()
    x = torch.rand(4, 3, 96, 96, device='cuda')
    w0 = model.decoder.conv_out.weight.detach().clone()
    for i in range(2):
        loss = tr.run_step(x, i)
    assert torch.isfinite(loss).all() and torch.isfinite(torch.as_tensor(model.logged['train/perc_loss'])).all()
    assert float(model.logged['train/perc_loss']) > 0 and not torch.equal(model.decoder.conv_out.weight, w0)


def test_down_conv_small_resolution_tensor_core_route(V):
    """the 8 -> 4 discriminator block: FIR output 9 x 9, full-resolution tcgen05 conv on a sub-tile-sized image + decimation."""
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.discriminator import Conv2dLayer, setup_filter
    torch.manual_seed(7)
    n, ci, co, h = 4, 128, 128, 8
    layer = Conv2dLayer(ci, co, kernel_size=3, activation='lrelu', down=2)
    with torch.no_grad():
        layer.weight.copy_(r16(layer.weight)); layer.bias.copy_(torch.randn(co) * 0.1)
    x = r16(torch.randn(n, ci, h, h)); go = r16(torch.randn(n, co, h // 2, h // 2))
    f = setup_filter()
    xo = x.clone().requires_grad_(); wo = layer.weight.detach().clone().requires_grad_(); bo = layer.bias.detach().clone().requires_grad_()
    xb = F.conv2d(F.pad(xo, [2, 2, 2, 2]), f[None, None].repeat(ci, 1, 1, 1), groups=ci)
    xb = xb + (r16(xb) - xb).detach()
    wq = wo * layer.weight_gain
    wq = wq + (r16(wq) - wq).detach()
    yo = F.leaky_relu(F.conv2d(xb, wq, bo, stride=2), 0.2) * np.sqrt(2)
    yo.backward(go)
    layer = layer.cuda()
    xg = cl(x).bfloat16().requires_grad_()
    yg = layer(xg)
    assert yg.shape == yo.shape
    yg.backward(cl(go).bfloat16())
    assert C.rel_err(yg.float(), yo) < 5e-3 and C.rel_err(xg.grad.float(), xo.grad) < 1e-2
    assert C.rel_err(layer.weight.grad, wo.grad) < 1e-2 and C.rel_err(layer.bias.grad, bo.grad) < 1e-2
