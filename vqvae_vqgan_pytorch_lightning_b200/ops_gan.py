"""autograd bindings of the loss-head kernels (csrc/gan_ops.cu): FIR resampling, max-pool, z-score, LPIPS taps,
minibatch stddev, NHWC->flat-NCHW for the discriminator's fully connected layers."""
from __future__ import annotations

import torch

from . import lib
from .lib import call, dt, ptr, stream
from .ops import as_nhwc, empty_nhwc


class Fir4Fn(torch.autograd.Function):
    """upfirdn2d(x, f=[1,3,3,1]^2/64, up=1, down=down, padding=pad)  (ops/upfirdn2d.py:120-208)."""

    @staticmethod
    def forward(ctx, x, pad, down):
        x = as_nhwc(x)
        n, c, h, w = x.shape
        oh, ow = (h + 2 * pad - 4) // down + 1, (w + 2 * pad - 4) // down + 1
        y = empty_nhwc(n, c, oh, ow, x.dtype, x.device)
        call('vqb_fir4_fwd', ptr(x), ptr(y), dt(x), n, h, w, c, pad, down, stream())
        ctx.cfg = (n, c, h, w, pad, down, x.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, pad, down, dtype = ctx.cfg
        return Fir4AdjointFn.apply(dy, h, w, pad, down, dtype), None, None


class Fir4AdjointFn(torch.autograd.Function):
    """adjoint of Fir4Fn (a linear map), itself an autograd function so that the graph of a backward pass can be recorded
    (R1 penalty): its backward is Fir4Fn again."""

    @staticmethod
    def forward(ctx, dy, h, w, pad, down, dtype):
        dy = as_nhwc(dy, dtype)
        n, c = dy.shape[0], dy.shape[1]
        dx = empty_nhwc(n, c, h, w, dtype, dy.device)
        call('vqb_fir4_bwd', ptr(dy), ptr(dx), dt(dy), n, h, w, c, pad, down, stream())
        ctx.cfg = (pad, down)
        return dx

    @staticmethod
    def backward(ctx, ddx):
        pad, down = ctx.cfg
        return Fir4Fn.apply(ddx, pad, down), None, None, None, None, None


def fir4(x, pad, down=1):
    return Fir4Fn.apply(x, pad, down)


class MaxPool2Fn(torch.autograd.Function):
    """nn.MaxPool2d(2, 2) of torchvision's VGG16 features."""

    @staticmethod
    def forward(ctx, x):
        x = as_nhwc(x)
        n, c, h, w = x.shape
        y = empty_nhwc(n, c, h // 2, w // 2, x.dtype, x.device)
        call('vqb_maxpool2_fwd', ptr(x), ptr(y), dt(x), n, h // 2, w // 2, c, stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, c, h, w = x.shape
        dy = as_nhwc(dy)
        dx = torch.zeros_like(x, dtype=dy.dtype, memory_format=torch.preserve_format) if (h % 2 or w % 2) else \
            torch.empty_like(x, dtype=dy.dtype, memory_format=torch.preserve_format)
        call('vqb_maxpool2_bwd', ptr(x), dt(x), ptr(dy), ptr(dx), dt(dy), n, h // 2, w // 2, c, stream())
        return dx


max_pool2 = MaxPool2Fn.apply


class MaxPool3s2Fn(torch.autograd.Function):
    """nn.MaxPool2d(kernel_size=3, stride=2) of torchvision's AlexNet features (overlapping windows)."""

    @staticmethod
    def forward(ctx, x):
        x = as_nhwc(x)
        n, c, h, w = x.shape
        y = empty_nhwc(n, c, (h - 3) // 2 + 1, (w - 3) // 2 + 1, x.dtype, x.device)
        call('vqb_maxpool3s2_fwd', ptr(x), ptr(y), dt(x), n, h, w, c, stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, c, h, w = x.shape
        dy = as_nhwc(dy)
        dx = torch.empty_like(x, dtype=dy.dtype, memory_format=torch.preserve_format)
        call('vqb_maxpool3s2_bwd', ptr(x), dt(x), ptr(dy), ptr(dx), dt(dy), n, h, w, c, stream())
        return dx


max_pool3s2 = MaxPool3s2Fn.apply


class ChannelAffineFn(torch.autograd.Function):
    """y = x * scale[c] + shift[c] (BaseNet.z_score with scale = 1/std, shift = -mean/std)."""

    @staticmethod
    def forward(ctx, x, scale, shift, out_dtype):
        x = as_nhwc(x)
        n, c, h, w = x.shape
        y = empty_nhwc(n, c, h, w, out_dtype or x.dtype, x.device)
        call('vqb_channel_affine', ptr(x), dt(x), ptr(y), dt(y), ptr(scale), ptr(shift), n * h * w, c, stream())
        ctx.save_for_backward(scale)
        ctx.in_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        dy = as_nhwc(dy)
        n, c, h, w = dy.shape
        dx = empty_nhwc(n, c, h, w, ctx.in_dtype, dy.device)
        zero = torch.zeros_like(scale)
        call('vqb_channel_affine', ptr(dy), dt(dy), ptr(dx), dt(dx), ptr(scale), ptr(zero), n * h * w, c, stream())
        return dx, None, None, None


def channel_affine(x, scale, shift, out_dtype=None):
    return ChannelAffineFn.apply(x, scale, shift, out_dtype)


class LpipsTapFn(torch.autograd.Function):
    """mean over batch of the spatial mean of lin((normalize(fx) - normalize(fy))^2)  (lpips.py:33-37) for ONE tap;
    differentiable w.r.t. fy (the reconstruction branch) only."""

    @staticmethod
    def forward(ctx, fx, fy, w):
        fx, fy = as_nhwc(fx), as_nhwc(fy)
        if fx.dtype != fy.dtype:
            fx = as_nhwc(fx, fy.dtype)
        n, c, h, wd = fy.shape
        out = torch.zeros(1, dtype=torch.float64, device=fy.device)
        wv = w.detach().reshape(-1).float().contiguous()
        call('vqb_lpips_tap_fwd', ptr(fx), ptr(fy), dt(fy), ptr(wv), ptr(out), n * h * wd, c, stream())
        ctx.save_for_backward(fx, fy, wv)
        return (out[0] / (n * h * wd)).float()

    @staticmethod
    def backward(ctx, g):
        fx, fy, wv = ctx.saved_tensors
        n, c, h, wd = fy.shape
        up = g.reshape(1).float().contiguous()
        dfy = torch.empty_like(fy, memory_format=torch.preserve_format)
        call('vqb_lpips_tap_bwd', ptr(fx), ptr(fy), dt(fy), ptr(wv), ptr(up), 1.0 / (n * h * wd), ptr(dfy), dt(dfy), n * h * wd, c,
             stream())
        return None, dfy, None


lpips_tap = LpipsTapFn.apply


class MbstdFn(torch.autograd.Function):
    """MinibatchStdLayer(group_size, num_channels=1)  (discriminator.py:277-293)."""

    @staticmethod
    def forward(ctx, x, group_size):
        x = as_nhwc(x)
        n, c, h, w = x.shape
        g = min(group_size, n) if group_size is not None else n
        if n % g:
            raise lib.VQBError(f'minibatch-stddev needs the batch ({n}) to be a multiple of the group size ({g})')
        y = empty_nhwc(n, c + 1, h, w, x.dtype, x.device)
        stat = torch.empty(n // g, dtype=torch.float32, device=x.device)
        call('vqb_mbstd_fwd', ptr(x), ptr(y), ptr(stat), dt(x), n, g, h * w, c, stream())
        ctx.save_for_backward(x)
        ctx.g = g
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, c, h, w = x.shape
        if torch.is_grad_enabled():
            # the graph of this backward pass is recorded (R1 penalty): the group standard deviation is the one non-linear,
            # non-piecewise-linear op of the discriminator, so its backward depends on x itself.  The tensor is tiny
            # (N x 512 x 4 x 4): restate the statistic with differentiable tensor ops and let the tape differentiate it.
            with torch.enable_grad():
                if not x.requires_grad:
                    raise lib.VQBError('minibatch-stddev: the saved input lost its autograd history (layout/dtype copy in forward)')
                y = _mbstd_composite(x.float(), ctx.g)
                (dx,) = torch.autograd.grad(y, x, dy.float(), create_graph=True)
            return dx.to(dy.dtype), None
        dy = as_nhwc(dy)
        dx = torch.empty_like(x, dtype=dy.dtype, memory_format=torch.preserve_format)
        call('vqb_mbstd_bwd', ptr(x), dt(x), ptr(dy), ptr(dx), dt(dy), n, ctx.g, h * w, c, stream())
        return dx, None


def _mbstd_composite(x, g):
    """discriminator.py:277-293 in tensor ops (fp32): used only for the second-order term of the R1 penalty."""
    n, c, h, w = x.shape
    y = x.reshape(g, n // g, c, h, w)
    y = y - y.mean(dim=0)
    y = (y.square().mean(dim=0) + 1e-8).sqrt()
    y = y.mean(dim=[1, 2, 3]).reshape(-1, 1, 1, 1).repeat(g, 1, h, w)
    return torch.cat([x, y], dim=1)


def mbstd(x, group_size=4):
    return MbstdFn.apply(x, group_size)


class FlattenNCHWFn(torch.autograd.Function):
    """x.flatten(1) of an NCHW tensor (discriminator.py:347) for a channels-last input: NHWC -> [N, C*H*W] in (c,h,w) order,
    returned as a channels-last [N, C*H*W, 1, 1] fp32 tensor so that the FC layer is a 1x1 implicit GEMM."""

    @staticmethod
    def forward(ctx, x):
        x = as_nhwc(x)
        n, c, h, w = x.shape
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        call('vqb_nhwc_to_nchw', ptr(x), dt(x), ptr(out), n, c, h, w, 1.0, 0.0, 0, 0.0, 0.0, stream())
        ctx.cfg = (n, c, h, w, x.dtype)
        return out.reshape(n, c * h * w, 1, 1)

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, dtype = ctx.cfg
        return UnflattenNCHWFn.apply(dy, c, h, w, dtype)


class UnflattenNCHWFn(torch.autograd.Function):
    """inverse (= adjoint) of FlattenNCHWFn: flat (c,h,w)-ordered [N, C*H*W, 1, 1] -> channels-last [N, C, H, W]."""

    @staticmethod
    def forward(ctx, dy, c, h, w, dtype):
        n = dy.shape[0]
        g = dy.reshape(n, c, h, w).float().contiguous()
        dx = empty_nhwc(n, c, h, w, dtype, dy.device)
        call('vqb_nchw_to_nhwc', ptr(g), ptr(dx), dt(dx), n, c, h, w, 0, 0.0, 0.0, 0.0, 1.0, stream())
        return dx

    @staticmethod
    def backward(ctx, ddx):
        return FlattenNCHWFn.apply(ddx), None, None, None, None


flatten_nchw = FlattenNCHWFn.apply


class Decimate2Fn(torch.autograd.Function):
    """y[..., oh, ow] = x[..., 2*oh+off, 2*ow+off]; backward = zero-insertion.  Together with a stride-1 'same' convolution at
    full resolution this reproduces the discriminator's unpadded stride-2 3x3 convolution (conv2d_resample.py:119-122)."""

    @staticmethod
    def forward(ctx, x, oh, ow, off):
        x = as_nhwc(x)
        n, c, h, w = x.shape
        y = empty_nhwc(n, c, oh, ow, x.dtype, x.device)
        call('vqb_decimate2', ptr(x), ptr(y), dt(x), n, h, w, c, oh, ow, off, stream())
        ctx.cfg = (n, c, h, w, oh, ow, off, x.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, oh, ow, off, dtype = ctx.cfg
        return ZeroUpsample2Fn.apply(dy, h, w, off, dtype), None, None, None


class ZeroUpsample2Fn(torch.autograd.Function):
    """adjoint of Decimate2Fn; its backward is Decimate2Fn (needed when the graph of a backward pass is recorded)."""

    @staticmethod
    def forward(ctx, dy, h, w, off, dtype):
        dy = as_nhwc(dy, dtype)
        n, c, oh, ow = dy.shape
        dx = empty_nhwc(n, c, h, w, dtype, dy.device)
        call('vqb_zero_upsample2', ptr(dy), ptr(dx), dt(dy), n, h, w, c, oh, ow, off, stream())
        ctx.cfg = (oh, ow, off)
        return dx

    @staticmethod
    def backward(ctx, ddx):
        oh, ow, off = ctx.cfg
        return Decimate2Fn.apply(ddx, oh, ow, off), None, None, None, None


def decimate2(x, oh, ow, off):
    return Decimate2Fn.apply(x, oh, ow, off)


# ------------------------------------------------------------------------------------------------------
# the discriminator's down-sampling 3x3 convolution (conv2d_resample.py:119-122) in space-to-depth form
# ------------------------------------------------------------------------------------------------------
def _s2d_packed(weight: torch.Tensor, w_scale: float):
    """Kernel layouts of a [co, c, 3, 3] weight for the 2x2-tap form: with t = 2a + dy (row) and u = 2b + dx (column) of the kernel
    zero-padded to 4x4, W2[co][a][b][(dy,dx,c)] = w[co][c][t][u].  Returns (forward K-major [co][(a,b),(dy,dx,c)], dgrad K-major
    [(dy,dx,c)][(1-a,1-b),co]) as persistent bf16 buffers, refreshed once per weights epoch."""
    from . import ops
    cache = getattr(weight, '_vqb_s2d', None)
    if cache is not None and cache['epoch'] == ops._weights_epoch and cache['scale'] == w_scale and cache['ptr'] == weight.data_ptr():
        return cache['fwd'], cache['dgrad']
    co, c = weight.shape[0], weight.shape[1]
    w4 = torch.nn.functional.pad(weight.detach().float() * w_scale, (0, 1, 0, 1)).view(co, c, 2, 2, 2, 2)      # [co, c, a, dy, b, dx]
    w2 = w4.permute(0, 2, 4, 3, 5, 1)                                                                         # [co, a, b, dy, dx, c]
    fwd = w2.reshape(co, 16 * c)
    dgr = w2.flip(1, 2).permute(3, 4, 5, 1, 2, 0).reshape(4 * c, 4 * co)
    if cache is None or cache['fwd'].numel() != fwd.numel() or cache['fwd'].device != weight.device:
        cache = {'fwd': torch.empty(fwd.shape, dtype=torch.bfloat16, device=weight.device),
                 'dgrad': torch.empty(dgr.shape, dtype=torch.bfloat16, device=weight.device)}
        weight._vqb_s2d = cache
    cache['fwd'].copy_(fwd); cache['dgrad'].copy_(dgr)
    cache.update(epoch=ops._weights_epoch, scale=w_scale, ptr=weight.data_ptr())
    return cache['fwd'], cache['dgrad']


def down2_conv3x3_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    from . import ops
    n, c, h, w = x.shape
    co = weight.shape[0]
    if ops.get_precision().name != 'fast' or tuple(weight.shape[2:]) != (3, 3) or h % 2 or w % 2 or c % 16 or co % 128:
        return False
    L = lib.load()
    h2, w2 = h // 2 + 1, w // 2 + 1
    return bool(L.vqb_conv2d_sub_supported(n, h2, w2, h // 2, w // 2, 4 * c, co, 2, 0) and
                L.vqb_conv2d_sub_supported(n, h // 2, w // 2, h2, w2, co, 4 * c, 2, -1))


class Down2Conv3x3Fn(torch.autograd.Function):
    """act(conv2d(FIR(x, pad 2), w * w_scale, stride 2) + bias) * gain -- conv2d_resample(down=2, padding=1) with the [1,3,3,1]
    filter (conv2d_resample.py:119-122) followed by bias_act -- without the full-resolution detour:
      z' = s2d(FIR(x))  [N, H/2+1, W/2+1, 4C]   one pass (vqb_fir4_s2d), then
      y[o] = sum_{a,b<2} W2[a][b] z'[o+a, o+b]  a 2x2-tap stride-1 tensor-core convolution over 4C channels (vqb_conv2d_fwd_sub):
    16C MACs per output instead of the 36C of a stride-1 3x3 evaluation at full resolution that is then decimated.  First order
    only: while the graph of a backward pass is recorded (R1) the layer takes the twice-differentiable route."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, alpha, gain, w_scale):
        from . import ops
        x = as_nhwc(x, torch.bfloat16)
        n, c, h, w = x.shape
        co = weight.shape[0]
        hz, wz, h2, w2, oh, ow = h + 1, w + 1, h // 2 + 1, w // 2 + 1, h // 2, w // 2
        zs = empty_nhwc(n, 4 * c, h2, w2, torch.bfloat16, x.device)
        call('vqb_fir4_s2d', ptr(x), ptr(zs), lib.BF16, n, h, w, c, hz, wz, 2, 0, 1, stream())
        wf, _ = _s2d_packed(weight, w_scale)
        b = bias.detach().reshape(-1).float().contiguous() if bias is not None else None
        y = empty_nhwc(n, co, oh, ow, torch.bfloat16, x.device)
        call('vqb_conv2d_fwd_sub', ptr(zs), ptr(wf), ptr(b), None, ptr(y), lib.BF16, n, h2, w2, oh, ow, 4 * c, co, 2, 0, act, alpha, gain,
             stream())
        ctx.save_for_backward(zs, weight, y if act != lib.ACT_NONE else None)
        ctx.cfg = (n, c, h, w, co, act, alpha, gain, w_scale, bias is not None)
        ctx.params = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        if torch.is_grad_enabled():
            raise lib.VQBError('Down2Conv3x3Fn is first-order only: run the discriminator under ops_gan.second_order() when the graph of '
                               'a backward pass is recorded (R1 penalty)')
        zs, weight, y = ctx.saved_tensors
        n, c, h, w, co, act, alpha, gain, w_scale, has_bias = ctx.cfg
        hz, wz, h2, w2, oh, ow = h + 1, w + 1, h // 2 + 1, w // 2 + 1, h // 2, w // 2
        dy = as_nhwc(dy)
        db = None
        want_db = has_bias and ctx.needs_input_grad[2]
        if act != lib.ACT_NONE:
            db_buf = ops.zero_arena.zeros(co, torch.float32, dy.device) if want_db else None
            dy = ops.ActBwdFn.apply(dy, y.detach(), act, alpha, gain, torch.bfloat16, db_buf)
            if want_db and ops.ActBwdFn.last_db_done:
                db = db_buf
        elif gain != 1.0:
            raise lib.VQBError('gain != 1 requires an activation epilogue')
        dy = as_nhwc(dy, torch.bfloat16)
        if want_db and db is None:
            db = ops._colsum(dy, n * oh * ow, co)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            _, wd = _s2d_packed(weight, w_scale)
            dzs = empty_nhwc(n, 4 * c, h2, w2, torch.bfloat16, dy.device)
            call('vqb_conv2d_fwd_sub', ptr(dy), ptr(wd), None, None, ptr(dzs), lib.BF16, n, oh, ow, h2, w2, co, 4 * c, 2, -1, lib.ACT_NONE, 0.0,
                 1.0, stream())
            dx = empty_nhwc(n, c, h, w, torch.bfloat16, dy.device)
            # adjoint of FIR(pad 2): the same filter with pad 1 on the (H+1) x (W+1) gradient, read from its space-to-depth layout
            call('vqb_fir4_s2d', ptr(dzs), ptr(dx), lib.BF16, n, hz, wz, c, h, w, 1, 1, 0, stream())
        if ctx.needs_input_grad[1] and not ops._no_weight_grad:
            dwp = torch.zeros(16 * c * co, dtype=torch.float32, device=dy.device)
            call('vqb_conv2d_wgrad_sub', ptr(zs), ptr(dy), ptr(dwp), n, h2, w2, oh, ow, 4 * c, co, 2, 0, stream())
            # dwp [(a,b)][(dy,dx,c)][co] -> dw [co][c][2a+dy][2b+dx] (the zero-padding row / column of the 4x4 frame is dropped)
            dw = dwp.view(2, 2, 2, 2, c, co).permute(5, 4, 0, 2, 1, 3).reshape(co, c, 4, 4)[:, :, :3, :3]
            dw = (dw * w_scale) if w_scale != 1.0 else dw.contiguous()
        return dx, dw, db, None, None, None, None


_second_order = 0


class second_order:
    """Context: the forward passes inside will be differentiated TWICE (R1: autograd.grad(..., create_graph=True), loss.py:98-112),
    so layers with a first-order-only fast route take their twice-differentiable one."""

    def __enter__(self):
        global _second_order
        _second_order += 1

    def __exit__(self, *a):
        global _second_order
        _second_order -= 1


def in_second_order() -> bool:
    return _second_order > 0


def down2_conv3x3(x, weight, bias, act, alpha, gain, w_scale):
    return Down2Conv3x3Fn.apply(x, weight, bias, act, alpha, gain, w_scale)
