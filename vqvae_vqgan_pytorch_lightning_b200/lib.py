"""ctypes binding of libvqgan_b200.so (the C ABI declared in include/vqgan_b200.h).

The product path FAILS LOUDLY when the shared library is missing or a call returns an error: there is no CPU or
eager-PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# VQB_LIB: another build of the same library (A/B measurements of one kernel on one box, tools/); default: the in-tree build
LIB_PATH = os.environ.get('VQB_LIB') or os.path.join(HERE, 'libvqgan_b200.so')

F32, BF16 = 0, 1
ACT_NONE, ACT_TANH, ACT_SILU, ACT_LRELU, ACT_RELU = 0, 1, 2, 3, 4

_p, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/vqgan_b200.h one to one (tests/test_abi.py checks the header)
SIGNATURES = {
    'vqb_last_error': (C.c_char_p, []),
    'vqb_version': (C.c_char_p, []),
    'vqb_device_supports_tcgen05': (_i, []),
    'vqb_nchw_to_nhwc': (_i, [_p, _p, _i, _i64, _i64, _i64, _i64, _i, _f, _f, _f, _f, _p]),
    'vqb_crop_flip_normalize': (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    'vqb_nhwc_to_nchw': (_i, [_p, _i, _p, _i64, _i64, _i64, _i64, _f, _f, _i, _f, _f, _p]),
    'vqb_split_hi_lo': (_i, [_p, _p, _i64, _i, _p]),
    'vqb_convert': (_i, [_p, _i, _p, _i, _i64, _p]),
    'vqb_pack_conv_weight': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _f, _p]),
    'vqb_pack_desc_bytes': (_sz, []),
    'vqb_pack_conv_weights_batched': (_i, [_p, _i, _i64, _p]),
    'vqb_im2col3x3_narrow': (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_unpack_conv_wgrad': (_i, [_p, _p, _i, _i, _i, _i, _f, _p]),
    'vqb_unpack_conv_wgrad_acc': (_i, [_p, _p, _i, _i, _i, _i, _f, _i, _i, _p]),
    'vqb_conv2d_fwd': (_i, [_i, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    'vqb_conv2d_fwd_gn': (_i, [_i, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p, _i, _p]),
    'vqb_conv2d_fwd_gn_supported': (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    'vqb_conv2d_wgrad': (_i, [_i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_set_halo_mode': (None, [_i]),
    'vqb_conv2d_dgrad': (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_fir4_fwd': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_fir4_bwd': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_decimate2': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_zero_upsample2': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_maxpool2_fwd': (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    'vqb_maxpool2_bwd': (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    'vqb_maxpool3s2_fwd': (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    'vqb_maxpool3s2_bwd': (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    'vqb_channel_affine': (_i, [_p, _i, _p, _i, _p, _p, _i64, _i, _p]),
    'vqb_lpips_tap_fwd': (_i, [_p, _p, _i, _p, _p, _i64, _i, _p]),
    'vqb_lpips_tap_bwd': (_i, [_p, _p, _i, _p, _p, _f, _p, _i, _i64, _i, _p]),
    'vqb_mbstd_fwd': (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    'vqb_mbstd_bwd': (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    'vqb_colsum': (_i, [_p, _i, _p, _i64, _i, _p]),
    'vqb_ssim_sums': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p]),
    'vqb_gn_stats': (_i, [_p, _i, _p, _i, _i, _i, _i, _p]),
    'vqb_gn_finalize': (_i, [_p, _p, _i, _i, _i, _i, _f, _p]),
    'vqb_gn_apply': (_i, [_p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_gn_bwd_reduce': (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    'vqb_gn_bwd_finalize': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    'vqb_gn_bwd_apply': (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    # x, xdt, sums, gamma, beta, y, ydt, stats_out, N, HW, C, G, eps, act, stream
    'vqb_gn_apply_sums': (_i, [_p, _i, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _f, _i, _p]),
    # x, xdt, dy, dydt, stats, gamma, beta, part, add, dx, dxdt, dgamma, dbeta, acc, N, HW, C, G, act, stream
    'vqb_gn_bwd_apply_part': (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    # x, wp, bias, residual, y, ydt, N, Hx, Wx, H, W, Ci, Co, T, off, act, alpha, gain, stream
    'vqb_conv2d_fwd_sub': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    'vqb_conv2d_wgrad_sub': (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    # x, wp, bias, y, ydt, N, H, W, Ci, Co, act, alpha, gain, stream
    'vqb_conv2d_fwd_narrowout': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    # x, xdt, wp, bias, residual, y, ydt, N, H, W, Ci, Co, act, alpha, gain, stream
    'vqb_conv2d_fwd_narrowin': (_i, [_p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    'vqb_conv2d_wgrad_narrow': (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    'vqb_conv2d_sub_supported': (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i]),
    # x, y, dtype, N, H, W, C, OH, OW, pad, in_s2d, out_s2d, stream
    'vqb_fir4_s2d': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_gn_bwd_fused_supported': (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    # x, dy, stats, gamma, beta, part, counters, add, dx, dgamma, dbeta, acc, N, HW, C, G, act, max_ctas, stream
    'vqb_gn_bwd_fused': (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    'vqb_down2': (_i, [_p, _p, _i, _i, _i, _i, _i, _f, _p]),
    'vqb_up2': (_i, [_p, _p, _i, _i, _i, _i, _i, _f, _p]),
    'vqb_diff_sums': (_i, [_p, _i, _p, _i, _p, _i64, _p]),
    'vqb_diff_grad': (_i, [_p, _i, _p, _i, _p, _i, _f, _f, _p, _i, _i64, _p]),
    'vqb_act_bwd_from_output': (_i, [_p, _i, _p, _i, _p, _i, _i, _f, _f, _i64, _p]),
    'vqb_act_bwd_bias': (_i, [_p, _p, _p, _i, _i, _f, _f, _i64, _i, _p, _p]),
    'vqb_vq_workspace_bytes': (_sz, [_i64, _i, _i]),
    'vqb_vq_assign': (_i, [_p, _p, _i, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _sz, _p]),
    'vqb_vq_tc_workspace_bytes': (_sz, [_i64, _i, _i]),
    'vqb_vq_assign_tc': (_i, [_p, _p, _i, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _sz, _p, _p]),
    'vqb_vq_prep_codebook': (_i, [_p, _p, _p, _i, _i, _p]),
    'vqb_vq_fused': (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _p]),
    'vqb_vq_fused_set_trace': (None, [_p]),
    'vqb_vq_ema_update_prep': (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _f, _f, _p]),
    'vqb_vq_ema_update': (_i, [_p, _p, _p, _p, _p, _i, _i, _f, _f, _f, _p]),
    'vqb_vq_backward': (_i, [_p, _p, _p, _p, _p, _f, _f, _p, _p, _i64, _i, _i, _p]),
    'vqb_vq_gather': (_i, [_p, _p, _p, _i64, _i, _i, _p]),
    'vqb_row_sqnorm': (_i, [_p, _p, _i64, _i, _p]),
    'vqb_vq_entropy_rows': (_i, [_p, _p, _p, _f, _p, _p, _i64, _i, _i, _i, _p]),
    'vqb_vq_colsum_exp': (_i, [_p, _p, _i64, _i, _p]),
    'vqb_vq_entropy_finalize': (_i, [_p, _p, _f, _p, _i64, _i, _p]),
    'vqb_vq_entropy_bwd_rows': (_i, [_p, _p, _p, _f, _f, _i64, _i, _p, _p]),
    'vqb_vq_entropy_combine_dcb': (_i, [_p, _p, _p, _p, _i, _i, _p]),
    'vqb_gumbel_rows_fwd': (_i, [_p, _p, _f, _i, _p, _p, _p, _i64, _i, _p]),
    'vqb_gumbel_rows_bwd': (_i, [_p, _p, _f, _p, _p, _f, _p, _i64, _i, _p]),
    'vqb_gumbel_rows_fwd_dev': (_i, [_p, _p, _p, _i, _p, _p, _p, _i64, _i, _p]),
    'vqb_gumbel_rows_bwd_dev': (_i, [_p, _p, _p, _p, _p, _f, _p, _i64, _i, _p]),
    'vqb_adamw': (_i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i, _f, _p]),
    'vqb_adamw_dev': (_i, [_p, _p, _p, _p, _i64, _p, _f, _f, _f, _f, _f, _p]),
}

_lib: Optional[C.CDLL] = None
launch_count = 0          # kernels-entry-point calls issued (bench.py reports this as gpu_launches)


class VQBError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (building is __graft_entry__.build()'s job, not an import side effect)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VQBError(f'{LIB_PATH} is missing: run `python -m vqvae_vqgan_pytorch_lightning_b200.build` '
                       f'(or __graft_entry__.build()); this package has no CPU / eager fallback')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise VQBError(f'unsupported dtype {t.dtype}')


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise VQBError('libvqgan_b200 kernels take CUDA tensors only (no CPU fallback)')
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class KernelTimer:
    """Optional per-entry-point device timing (CUDA events on the launching stream, resolved after a sync).
    bench.py uses it to measure the dominant kernel's average launch duration inside the timed region."""

    def __init__(self, names):
        self.names = set(names)
        self.records = []          # (name, args, start_event, end_event)

    def summary(self):
        out = {}
        for name, args, e0, e1 in self.records:
            d = out.setdefault(name, {'calls': 0, 'ms': 0.0, 'args': []})
            d['calls'] += 1
            d['ms'] += e0.elapsed_time(e1)
            d['args'].append(args)
        return out


timer: Optional[KernelTimer] = None


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise on a non-zero status."""
    global launch_count
    lib = load()
    if timer is not None and name in timer.names:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        timer.records.append((name, args, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    launch_count += 1
    if rc != 0:
        raise VQBError(f'{name} failed ({rc}): {lib.vqb_last_error().decode()}')
