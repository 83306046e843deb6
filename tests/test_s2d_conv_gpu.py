"""GPU: the discriminator's down-sampling convolution in space-to-depth form (stylegan2_discriminator/ops/conv2d_resample.py:119-122,
upfirdn2d.py:120-208) against plain-torch fp32 restatements on bf16-rounded operands:

* the sliding-window FIR kernel (fp32 / bf16, ragged sizes, several strips) and its space-to-depth output / input forms,
* T x T-tap sub-convolutions on the 3x3 halo kernels: forward (the three kernels: one-CTA, swapped-operand, CTA-pair), the dgrad form
  (off = -1, output larger than the input) and the weight gradient,
* the whole layer (FIR + stride-2 3x3 conv + bias + lrelu * gain), forward and backward, against the reference arithmetic."""
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


def fir_ref(x, pad, oh=None, ow=None):
    c = x.shape[1]
    f = torch.tensor([1., 3., 3., 1.], dtype=x.dtype, device=x.device); f = f.ger(f); f = f / f.sum()
    y = F.conv2d(F.pad(x, [pad] * 4), f[None, None].repeat(c, 1, 1, 1), groups=c)
    return y if oh is None else y[:, :, :oh, :ow]


def s2d(t):
    """[N,C,H,W] (H, W even) -> [N,(dy,dx,c),H/2,W/2]"""
    n, c, h, w = t.shape
    return t.view(n, c, h // 2, 2, w // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, h // 2, w // 2)


@pytest.mark.parametrize('dtype,n,c,h,w,pad', [(torch.float32, 2, 8, 17, 23, 2), (torch.float32, 1, 4, 70, 9, 1), (torch.bfloat16, 2, 64, 40, 36, 2),
                                               (torch.bfloat16, 1, 128, 33, 16, 1), (torch.float32, 2, 16, 6, 5, 0)])
def test_fir_strip_kernel(V, dtype, n, c, h, w, pad):
    from vqvae_vqgan_pytorch_lightning_b200 import ops_gan
    torch.manual_seed(0)
    x = torch.randn(n, c, h, w)
    x = r16(x) if dtype == torch.bfloat16 else x
    xo = x.clone().requires_grad_()
    y = fir_ref(xo, pad)
    go = torch.randn_like(y); go = r16(go) if dtype == torch.bfloat16 else go
    y.backward(go)
    xg = cl(x).to(dtype).requires_grad_()
    yg = ops_gan.fir4(xg, pad, 1)
    yg.backward(cl(go).to(dtype))
    tol = 4e-3 if dtype == torch.bfloat16 else 1e-6
    assert yg.shape == y.shape
    assert C.rel_err(yg.float(), y) < tol and C.rel_err(xg.grad.float(), xo.grad) < tol


@pytest.mark.parametrize('dtype,n,c,h,w', [(torch.float32, 2, 8, 18, 22), (torch.float32, 1, 4, 70, 10), (torch.bfloat16, 2, 64, 40, 36),
                                           (torch.bfloat16, 1, 128, 34, 16), (torch.float32, 2, 16, 6, 4), (torch.bfloat16, 3, 8, 64, 64)])
def test_fir_down2_strip_kernels(V, dtype, n, c, h, w):
    """the discriminator's skip path: upfirdn2d(x, f, down=2, padding=1) forward and adjoint on the strip / cp.async-ring kernels"""
    from vqvae_vqgan_pytorch_lightning_b200 import ops_gan
    torch.manual_seed(7)
    x = torch.randn(n, c, h, w)
    x = r16(x) if dtype == torch.bfloat16 else x
    xo = x.clone().requires_grad_()
    y = fir_ref(xo, 1)[:, :, ::2, ::2]
    go = torch.randn_like(y); go = r16(go) if dtype == torch.bfloat16 else go
    y.backward(go)
    xg = cl(x).to(dtype).requires_grad_()
    yg = ops_gan.fir4(xg, 1, 2)
    yg.backward(cl(go).to(dtype))
    tol = 4e-3 if dtype == torch.bfloat16 else 1e-6
    assert yg.shape == y.shape
    assert C.rel_err(yg.float(), y) < tol and C.rel_err(xg.grad.float(), xo.grad) < tol


@pytest.mark.parametrize('n,c,h,w', [(2, 16, 16, 16), (1, 64, 34, 20), (2, 8, 6, 10)])
def test_fir_s2d_forms(V, n, c, h, w):
    from vqvae_vqgan_pytorch_lightning_b200.lib import BF16, call, ptr, stream
    torch.manual_seed(1)
    x = r16(torch.randn(n, c, h, w))
    z = fir_ref(x, 2)                                                   # [n, c, h+1, w+1]
    zp = F.pad(z, [0, 1, 0, 1])                                          # the padding slot of the (h+2) x (w+2) physical tensor is zero
    ref = s2d(zp)
    xg = cl(x).bfloat16()
    h2, w2 = h // 2 + 1, w // 2 + 1
    zs = torch.full((n, 4 * c, h2, w2), float('nan'), dtype=torch.bfloat16, device='cuda').contiguous(memory_format=torch.channels_last)
    call('vqb_fir4_s2d', ptr(xg), ptr(zs), BF16, n, h, w, c, h + 1, w + 1, 2, 0, 1, stream())
    assert torch.isfinite(zs.float()).all()
    assert C.rel_err(zs.float(), ref) < 4e-3
    assert float(zs.float().cpu()[:, :, -1, :].view(n, 2, 2, c, w2)[:, 1].abs().max()) == 0.0      # dy = 1 of the last physical row: padding
    # adjoint: dz (logical (h+1) x (w+1), stored space-to-depth with garbage in the padding slot) -> dx = FIR(dz, pad 1)
    dz = r16(torch.randn(n, c, h + 1, w + 1))
    dzp = F.pad(dz, [0, 1, 0, 1], value=1e4)                             # the padding slot must be IGNORED
    dzs = cl(s2d(dzp)).bfloat16()
    dx = torch.empty((n, c, h, w), dtype=torch.bfloat16, device='cuda').contiguous(memory_format=torch.channels_last)
    call('vqb_fir4_s2d', ptr(dzs), ptr(dx), BF16, n, h + 1, w + 1, c, h, w, 1, 1, 0, stream())
    assert C.rel_err(dx.float(), fir_ref(dz, 1)) < 4e-3


def sub_ref(x, w2, off, oh, ow):
    """y[o] = sum_{a,b<T} w2[:, :, a, b] x[o + off + a, o + off + b], x zero outside; output oh x ow"""
    T = w2.shape[2]
    n, ci, hx, wx = x.shape
    lo = -off                                                            # zero rows before index 0
    xp = F.pad(x, [lo, max(0, ow + off + T - 1 - wx), lo, max(0, oh + off + T - 1 - hx)])
    return F.conv2d(xp, w2)[:, :, :oh, :ow]


@pytest.mark.parametrize('n,hx,wx,oh,ow,ci,co,off', [
    (2, 17, 9, 16, 8, 64, 64, 0),          # one-CTA halo kernel, smallest output
    (2, 33, 33, 32, 32, 128, 128, 0),      # swapped-operand kernel (128-channel tile, H >= 32)
    (3, 17, 17, 16, 16, 256, 256, 0),      # CTA-pair kernel
    (2, 16, 16, 17, 17, 256, 256, -1),     # dgrad form: output larger than the input
    (1, 32, 24, 33, 25, 128, 512, -1),
])
def test_sub_convolution_forward(V, n, hx, wx, oh, ow, ci, co, off):
    from vqvae_vqgan_pytorch_lightning_b200.lib import ACT_NONE, BF16, call, ptr, stream
    torch.manual_seed(2)
    x = r16(torch.randn(n, ci, hx, wx))
    w2 = r16(torch.randn(co, ci, 2, 2) / np.sqrt(4 * ci))
    ref = sub_ref(x, w2, off, oh, ow)
    wp = w2.permute(0, 2, 3, 1).reshape(co, 4 * ci).contiguous().cuda().bfloat16()          # [co][(a,b),ci]
    xg = cl(x).bfloat16()
    y = torch.empty((n, co, oh, ow), dtype=torch.float32, device='cuda').contiguous(memory_format=torch.channels_last)
    call('vqb_conv2d_fwd_sub', ptr(xg), ptr(wp), None, None, ptr(y), V.lib.F32, n, hx, wx, oh, ow, ci, co, 2, off, ACT_NONE, 0.0, 1.0, stream())
    assert C.rel_err(y, ref) < 1e-4                                       # fp32 output: accumulation order only


@pytest.mark.parametrize('n,hx,wx,oh,ow,ci,co', [(2, 17, 17, 16, 16, 64, 128), (2, 33, 17, 32, 16, 256, 256)])
def test_sub_convolution_wgrad(V, n, hx, wx, oh, ow, ci, co):
    from vqvae_vqgan_pytorch_lightning_b200.lib import call, ptr, stream
    torch.manual_seed(3)
    x = r16(torch.randn(n, ci, hx, wx))
    dy = r16(torch.randn(n, co, oh, ow))
    w2 = torch.zeros(co, ci, 2, 2, requires_grad=True)
    (sub_ref(x, w2, 0, oh, ow) * dy).sum().backward()
    dwp = torch.zeros(4 * ci * co, device='cuda')
    xg, dyg = cl(x).bfloat16(), cl(dy).bfloat16()                        # (named: a temporary's storage is recycled after ptr())
    call('vqb_conv2d_wgrad_sub', ptr(xg), ptr(dyg), ptr(dwp), n, hx, wx, oh, ow, ci, co, 2, 0, stream())
    got = dwp.view(2, 2, ci, co).permute(3, 2, 0, 1)
    assert C.rel_err(got, w2.grad) < 1e-4


@pytest.mark.parametrize('activation', ['linear', 'lrelu'])
@pytest.mark.parametrize('n,c,co,res', [(2, 64, 128, 32), (2, 128, 256, 64)])
def test_down2_layer_matches_reference_arithmetic(V, n, c, co, res, activation):
    from vqvae_vqgan_pytorch_lightning_b200 import ops_gan
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss import discriminator as D
    torch.manual_seed(4)
    layer = D.Conv2dLayer(c, co, kernel_size=3, activation=activation, down=2).cuda()
    with torch.no_grad():
        layer.bias.copy_(torch.randn(co) * 0.1)
    x = r16(torch.randn(n, c, res, res))
    go = r16(torch.randn(n, co, res // 2, res // 2))
    lrelu = activation == 'lrelu'
    gain = float(np.sqrt(0.5)) if lrelu else 1.0
    # reference arithmetic in fp32 (conv2d_resample down=2 + bias_act): FIR pad 2, stride-2 conv, bias [, lrelu(0.2) * sqrt(2) * gain]
    xo = x.clone().requires_grad_()
    wo = layer.weight.detach().cpu().clone().requires_grad_()
    bo = layer.bias.detach().cpu().clone().requires_grad_()
    yo = F.conv2d(fir_ref(xo, 2), wo * layer.weight_gain, stride=2) + bo.view(1, -1, 1, 1)
    if lrelu:
        yo = F.leaky_relu(yo, 0.2) * (np.sqrt(2) * gain)
    yo.backward(go)
    outs = {}
    for route in (True, False):
        D._S2D_ROUTE = route
        layer.zero_grad()
        xg = cl(x).bfloat16().requires_grad_()
        yg = layer(xg, gain=gain)
        yg.backward(cl(go).bfloat16())
        outs[route] = (yg.float().cpu(), xg.grad.float().cpu(), layer.weight.grad.cpu().clone(), layer.bias.grad.cpu().clone())
    D._S2D_ROUTE = True
    assert ops_gan.down2_conv3x3_supported(cl(x).bfloat16(), layer.weight)
    # linear: bf16 rounding of operands / intermediates only.  lrelu: the slope of a pre-activation within bf16 rounding of zero
    # flips against the fp32 reference (0.1 % flipped slopes put ~3 % L2 error on a gradient) -- flip-tolerant bar,
    # and the two routes (same rounding points) are held to each other
    gtol = 6e-2 if lrelu else 1.2e-2
    for route, (y, dx, dw, db) in outs.items():
        assert C.rel_err(y, yo) < 6e-3, route
        assert C.rel_err(dx, xo.grad) < gtol, (route, C.rel_err(dx, xo.grad))
        assert C.rel_err(dw, wo.grad) < gtol and C.rel_err(db, bo.grad) < gtol, route
    assert C.rel_err(outs[True][0], outs[False][0]) < 6e-3
    assert C.rel_err(outs[True][1], outs[False][1]) < gtol and C.rel_err(outs[True][2], outs[False][2]) < gtol
