"""GPU parity of the tcgen05/TMA implicit-GEMM convolution (bf16 operands, fp32 accumulate) against the CPU oracle.

Inputs are pre-rounded to bf16 so that the oracle (fp32 F.conv2d on the same rounded values) differs only by
accumulation order: fp32 outputs must agree to 1e-4 relative; bf16 outputs to bf16 rounding (4e-3)."""
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


CASES = [
    # n, h, w, ci, co, k, bias, res, act
    (2, 16, 16, 64, 64, 3, False, False, 0),
    (1, 16, 16, 128, 128, 3, True, True, 0),
    (3, 16, 16, 128, 256, 3, False, False, 0),
    (2, 32, 32, 256, 128, 3, False, True, 0),
    (2, 8, 8, 512, 256, 1, True, False, 0),
    (5, 4, 4, 128, 128, 3, False, False, 0),        # spatial smaller than a tile: several images per TMA box, N not a multiple
    (2, 12, 20, 64, 128, 3, False, False, 0),       # H, W not powers of two: partially out-of-bounds boxes
    (1, 64, 64, 128, 128, 3, False, False, 0),
    (2, 16, 16, 512, 512, 3, False, False, 0),      # two Co tiles, 72 k-steps (pipeline wrap-around)
    (2, 40, 20, 128, 128, 3, True, True, 0),        # swapped-operand kernel (M = co, N = 256 pixels), ragged 32x8 tiles
    (1, 33, 9, 256, 128, 3, False, True, 0),
    (2, 32, 32, 128, 3, 3, True, False, 1),         # narrow head (decoder.conv_out): UMMA N=16, tanh epilogue
]


@pytest.mark.parametrize('n,h,w,ci,co,k,bias,res,act', CASES)
def test_conv_tc_forward_fp32_out(V, n, h, w, ci, co, k, bias, res, act):
    torch.manual_seed(0)
    x = r16(torch.randn(n, ci, h, w))
    wt = r16(torch.randn(co, ci, k, k) / (ci * k * k) ** 0.5)
    b = torch.randn(co) if bias else None
    r = torch.randn(n, co, h, w) if res else None
    y = F.conv2d(x, wt, b, padding=k // 2)
    if act == 1:
        y = torch.tanh(y)
    if res:
        y = y + r
    yg = V.ops.conv2d(cl(x).bfloat16(), wt.cuda(), b.cuda() if bias else None, cl(r) if res else None, pad=k // 2,
                      act=act, out_dtype=torch.float32)
    assert yg.dtype == torch.float32
    assert C.rel_err(yg, y) < 1e-4
    assert C.max_rel(yg, y) < 5e-3


@pytest.mark.parametrize('n,h,w,ci,co,k,bias,res,act', CASES)
def test_conv_tc_backward(V, n, h, w, ci, co, k, bias, res, act):
    torch.manual_seed(1)
    x = r16(torch.randn(n, ci, h, w))
    wt = r16(torch.randn(co, ci, k, k) / (ci * k * k) ** 0.5)
    go = r16(torch.randn(n, co, h, w))
    xo, wo = x.clone().requires_grad_(), wt.clone().requires_grad_()
    F.conv2d(xo, wo, None, padding=k // 2).backward(go)
    xg, wg = cl(x).bfloat16().requires_grad_(), wt.cuda().requires_grad_()
    yg = V.ops.conv2d(xg, wg, None, None, pad=k // 2)
    assert yg.dtype == torch.bfloat16
    yg.backward(cl(go).bfloat16())
    assert C.rel_err(xg.grad.float(), xo.grad) < 4e-3          # dgrad output is stored as bf16
    assert C.rel_err(wg.grad, wo.grad) < 1e-4                  # wgrad accumulates and stores fp32


def test_rgb_heads_on_tensor_cores(V):
    """encoder.conv_in (3 -> 128) and decoder.conv_out (128 -> 3, tanh): forward / dgrad / wgrad through the im2col-64
    1x1 tcgen05 route, against fp32 torch on the same bf16-rounded operands."""
    torch.manual_seed(2)
    n, h, w = 2, 40, 24
    # --- narrow input
    x = r16(torch.rand(n, 3, h, w) * 2 - 1); wt = r16(torch.randn(128, 3, 3, 3) * 0.2); go = r16(torch.randn(n, 128, h, w))
    wo = wt.clone().requires_grad_()
    yo = F.conv2d(x, wo, None, padding=1); yo.backward(go)
    wg = wt.cuda().requires_grad_()
    yg = V.ops.conv2d(cl(x), wg, None, None, pad=1)                    # fp32 image input, bf16 output
    yg.backward(cl(go).bfloat16())
    assert yg.dtype == torch.bfloat16
    assert C.rel_err(yg.float(), yo) < 4e-3 and C.rel_err(wg.grad, wo.grad) < 1e-4
    # --- narrow output with bias + tanh
    x = r16(torch.randn(n, 128, h, w)); wt = r16(torch.randn(3, 128, 3, 3) * 0.03); b = torch.randn(3) * 0.1
    go = torch.randn(n, 3, h, w)
    xo, wo, bo = x.clone().requires_grad_(), wt.clone().requires_grad_(), b.clone().requires_grad_()
    yo = torch.tanh(F.conv2d(xo, wo, bo, padding=1)); yo.backward(go)
    xg, wg, bg = cl(x).bfloat16().requires_grad_(), wt.cuda().requires_grad_(), b.cuda().requires_grad_()
    yg = V.ops.conv2d(xg, wg, bg, None, pad=1, act=V.lib.ACT_TANH, out_dtype=torch.float32)
    yg.backward(cl(go))
    assert C.rel_err(yg, yo) < 1e-4
    assert C.rel_err(xg.grad.float(), xo.grad) < 8e-3              # dpre and dx are stored as bf16
    assert C.rel_err(wg.grad, wo.grad) < 5e-3 and C.rel_err(bg.grad, bo.grad) < 5e-3


@pytest.mark.parametrize('n,h,w,ci,co,k,bias,res,f32out', [
    (2, 64, 64, 128, 128, 3, True, True, False),     # swapped-operand kernel (conv_fwd_tc_halo_t), 4 channels per group, bf16 out
    (3, 40, 24, 128, 128, 3, False, False, True),    # same kernel, ragged tiles, fp32 out
    (2, 32, 32, 256, 256, 3, False, True, False),    # CTA-pair kernel (halo2), 8 channels per group
    (3, 16, 16, 256, 512, 3, True, False, False),    # CTA-pair kernel, 16 channels per group, odd number of pixel tiles
    (5, 64, 64, 256, 512, 3, False, False, False),   # two channel tiles per pixel tile, many images per CTA (flush on image change)
    (2, 16, 16, 128, 128, 3, False, True, False),    # one-CTA halo kernel (H < 32): NOT fused -> GroupNorm runs its own pass
    (2, 32, 16, 256, 256, 1, True, False, False),    # generic kernel (1x1): not fused either
])
def test_conv_epilogue_groupnorm_statistics(V, n, h, w, ci, co, k, bias, res, f32out):
    """vqb_conv2d_fwd_gn: the per-(image, group) sum / sum of squares of the convolution OUTPUT from the epilogue registers equal
    those of the stored tensor (fp32 output: to accumulation order; bf16 output: the statistics see the values before the
    bf16 rounding, 2^-9 relative per element with random sign), and GroupNorm fed with them equals GroupNorm on its own pass."""
    torch.manual_seed(11)
    x = r16(torch.randn(n, ci, h, w))
    wt = r16(torch.randn(co, ci, k, k) / (ci * k * k) ** 0.5)
    b = torch.randn(co) if bias else None
    r = torch.randn(n, co, h, w) if res else None
    out_dtype = torch.float32 if f32out else torch.bfloat16
    rg = cl(r).to(out_dtype) if res else None
    y = V.ops.conv2d(cl(x).bfloat16(), wt.cuda(), b.cuda() if bias else None, rg, pad=k // 2, out_dtype=out_dtype, gn_groups=32)
    import os
    fused = k == 3 and ((co % 256 == 0 and os.environ.get('VQB_GN_FUSE') == '2') or (co % 256 != 0 and h >= 32))
    assert hasattr(y, '_gn_sums') == fused, 'by default only the swapped-operand 3x3 kernel fuses the statistics (VQB_GN_FUSE=2: CTA pairs too)'
    if not fused:
        return
    sums, groups = y._gn_sums
    assert groups == 32 and sums.numel() == n * 32 * 2
    yf = y.float().reshape(n, 32, co // 32, h, w).double()
    ref = torch.stack([yf.sum(dim=(2, 3, 4)), (yf * yf).sum(dim=(2, 3, 4))], dim=-1).reshape(-1)
    tol = 1e-5 if f32out else 2e-3
    scale = ref.reshape(n, 32, 2)[..., 1].sqrt().reshape(n, 32, 1).expand(n, 32, 2).reshape(-1) * (co // 32 * h * w) ** 0.5   # |x|_2 * sqrt(count)
    got = sums.reshape(n, 32, 2)
    assert float(((got[..., 0].reshape(-1) - ref.reshape(n, 32, 2)[..., 0].reshape(-1).cuda()).abs() / scale.reshape(n, 32, 2)[..., 0].reshape(-1).cuda()).max()) < tol
    assert C.rel_err(got[..., 1], ref.reshape(n, 32, 2)[..., 1]) < tol
    # GroupNorm + SiLU with the fused statistics vs its own statistics pass on the same tensor
    gamma, beta = torch.randn(1, co, 1, 1).cuda(), torch.randn(1, co, 1, 1).cuda()
    o1 = V.ops.group_norm_act(y, gamma, beta)
    y2 = y.detach().clone()
    o2 = V.ops.group_norm_act(y2, gamma, beta)
    assert C.rel_err(o1.float(), o2.float()) < (1e-5 if f32out else 6e-3)


@pytest.mark.parametrize('n,ci,h,w,co,act', [(2, 128, 32, 32, 3, 'tanh'), (1, 128, 19, 21, 3, 'none'), (2, 128, 16, 16, 1, 'tanh'),
                                              (1, 256, 40, 28, 3, 'tanh'), (3, 128, 64, 64, 2, 'none')])
def test_narrow_output_head_partial_products(n, ci, h, w, co, act):
    """decoder.conv_out-like heads (Co <= 3, autoencoder.py:170,178) on the per-tap-partials kernel (vqb_conv2d_fwd_narrowout): forward
    against fp32 torch on bf16-rounded operands (fp32 output: accumulation order only), and the module path with its backward against
    the N = 16 halo kernel it replaces"""
    import torch.nn.functional as F
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    from vqvae_vqgan_pytorch_lightning_b200 import ops
    from vqvae_vqgan_pytorch_lightning_b200.lib import ACT_NONE, ACT_TANH
    pkg.lib.load()
    if not pkg.lib.load().vqb_device_supports_tcgen05():
        pytest.skip('needs sm_100')
    pkg.set_precision('fast')
    try:
        torch.manual_seed(11)
        x = torch.randn(n, ci, h, w).bfloat16().float()
        wt = (torch.randn(co, ci, 3, 3) / (3 * ci ** 0.5)).bfloat16().float()
        b = torch.randn(co) * 0.1
        ref = F.conv2d(x, wt, b, padding=1)
        ref = torch.tanh(ref) if act == 'tanh' else ref
        outs = {}
        for flag in (True, False):
            ops._narrowout = flag
            xg = x.cuda().contiguous(memory_format=torch.channels_last).bfloat16().requires_grad_()
            wg = torch.nn.Parameter(wt.cuda().clone())
            bg = torch.nn.Parameter(b.cuda().clone())
            y = ops.conv2d(xg, wg, bg, None, pad=1, act=ACT_TANH if act == 'tanh' else ACT_NONE, out_dtype=torch.float32)
            y.backward(torch.ones_like(y))
            outs[flag] = (y.detach().float().cpu(), xg.grad.float().cpu(), wg.grad.cpu().clone())
        err = float((outs[True][0] - ref).norm() / ref.norm())
        assert err < 1e-4, err
        assert float((outs[True][0] - outs[False][0]).norm() / ref.norm()) < 1e-4
        assert torch.equal(outs[True][1], outs[False][1]) or float((outs[True][1] - outs[False][1]).norm() / outs[False][1].norm()) < 1e-2
        assert float((outs[True][2] - outs[False][2]).norm() / outs[False][2].norm()) < 1e-2
    finally:
        ops._narrowout = None
        pkg.set_precision('strict')


@pytest.mark.parametrize('n,h,w,co,dtype,act', [(2, 32, 32, 128, torch.float32, 'none'), (1, 19, 21, 128, torch.bfloat16, 'none'),
                                                 (2, 16, 16, 64, torch.bfloat16, 'relu'), (3, 64, 48, 128, torch.float32, 'none'),
                                                 (1, 8, 8, 256, torch.float32, 'none')])
def test_narrow_input_head_in_kernel_im2col(n, h, w, co, dtype, act):
    """encoder.conv_in / VGG conv1_1-like heads (Ci = 3) with the A operand built inside the kernel (vqb_conv2d_fwd_narrowin) against
    fp32 torch on bf16-rounded operands, and against the im2col route it replaces (same rounding points: fp32 outputs equal to
    accumulation order)"""
    import torch.nn.functional as F
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    from vqvae_vqgan_pytorch_lightning_b200 import ops
    from vqvae_vqgan_pytorch_lightning_b200.lib import ACT_NONE, ACT_RELU
    pkg.lib.load()
    if not pkg.lib.load().vqb_device_supports_tcgen05():
        pytest.skip('needs sm_100')
    pkg.set_precision('fast')
    try:
        torch.manual_seed(12)
        x = torch.randn(n, 3, h, w)
        x16 = x.bfloat16().float()
        wt = (torch.randn(co, 3, 3, 3) / 27 ** 0.5).bfloat16().float()
        b = torch.randn(co) * 0.1
        ref = F.conv2d(x16, wt, b, padding=1)
        ref = torch.relu(ref) if act == 'relu' else ref
        outs = {}
        for flag in (True, False):
            ops._narrowin = flag
            xg = (x16 if dtype == torch.bfloat16 else x).cuda().contiguous(memory_format=torch.channels_last).to(dtype)
            wg = torch.nn.Parameter(wt.cuda().clone(), requires_grad=(co % 128 == 0))     # (64-wide heads ride this route only when frozen)
            bg = torch.nn.Parameter(b.cuda().clone())
            y = ops.conv2d(xg, wg, bg, None, pad=1, act=ACT_RELU if act == 'relu' else ACT_NONE, out_dtype=torch.float32)
            gw = None
            if wg.requires_grad:                      # weight gradient: vqb_conv2d_wgrad_narrow vs im2col + the generic 1x1 kernel
                torch.manual_seed(13)
                go = torch.randn(n, co, h, w).bfloat16().float()
                y.backward(go.cuda().contiguous(memory_format=torch.channels_last))
                gw = wg.grad.cpu().clone()
            outs[flag] = (y.detach().float().cpu(), gw)
        err = float((outs[True][0] - ref).norm() / ref.norm())
        assert err < 1e-4, err
        assert float((outs[True][0] - outs[False][0]).norm() / ref.norm()) < 1e-4
        if outs[True][1] is not None:
            wr = wt.clone().requires_grad_()
            out = F.conv2d(x16, wr, b, padding=1)
            torch.manual_seed(13)
            out.backward(torch.randn(n, co, h, w).bfloat16().float())
            assert float((outs[True][1] - wr.grad).norm() / wr.grad.norm()) < 2e-4, 'narrow weight gradient vs fp32 torch'
            assert float((outs[True][1] - outs[False][1]).norm() / wr.grad.norm()) < 2e-4
    finally:
        ops._narrowin = None
        pkg.set_precision('strict')
