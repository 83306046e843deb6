"""torch.autograd bindings of the libvqgan_b200 kernels.

PyTorch is used here for device memory, streams and the autograd tape only; every arithmetic operation on the
hot path is a call through the C ABI (lib.call).  Activations are channels-last tensors (logical [N,C,H,W],
physical NHWC) so the module surface keeps the reference's shapes while the kernels see NHWC.
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import lib
from .lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SILU, ACT_TANH, BF16, F32, call, dt, ptr, stream

CL = torch.channels_last


@dataclass
class Precision:
    """Numeric mode of the convolution stacks.

    strict: fp32 storage; convolutions on the tensor cores with SPLIT-PRECISION operands -- every fp32 value is split into bf16
            hi + lo (16 mantissa bits), three (or four) bf16 products per multiply accumulate in ONE fp32 TMEM accumulator,
            by the same tcgen05 kernels as the fast mode (their k loop wraps over the [hi | lo] halves) -- the parity path
            (<= 1e-4 rel. vs the fp32 oracle); shapes the tensor-core kernels do not take (RGB heads, strided or odd-channel
            discriminator layers, problems of fewer than 1024 output pixels) run on the fp32 SIMT implicit GEMM.  Default:
            three products (hi.hi + lo.hi + hi.lo, ~1.2e-5 per convolution: measured in tests/test_bench_shapes_gpu.py; the whole
            GPU parity suite holds its 1e-4 bars with it); VQB_STRICT_CONV=tc4 adds lo.lo (~4e-6, 4/3 slower), =simt forces the
            SIMT kernels.
    fast  : bf16 storage, tcgen05 bf16 x bf16 -> fp32 implicit GEMM wherever Ci and Co are multiples of 64
            (the reference itself trains with precision='16-mixed', vqvae/train.py:129); fp32 master weights,
            fp32 GroupNorm statistics, fp32 VQ, fp32 weight gradients.
    """
    name: str = 'strict'

    @property
    def act_dtype(self) -> torch.dtype:
        return torch.float32 if self.name == 'strict' else torch.bfloat16

    def conv_impl(self, ci: int, co: int, stride: int = 1, pixels: int = 1 << 30) -> int:
        """1 = tcgen05 implicit GEMM on bf16 operands (forward, and dgrad with ci/co swapped), 2 / 3 = the same kernels on
        split-precision operands (3 / 4 bf16 products per multiply), 0 = fp32 SIMT.  `pixels` = N*OH*OW of the problem."""
        if self.name == 'fast' and ci % 64 == 0 and (co % 64 == 0 or co <= 16) and stride == 1:
            return 1
        if self.name == 'strict' and _strict_conv() and ci % 64 == 0 and co % 64 == 0 and stride == 1 and pixels >= 1024:
            return _strict_conv()
        return 0

    def wgrad_impl(self, ci: int, co: int, stride: int = 1, pixels: int = 1 << 30) -> int:
        if self.name == 'fast' and ci % 64 == 0 and co % 128 == 0 and stride == 1:
            return 1
        if self.name == 'strict' and _strict_conv() and ci % 64 == 0 and co % 64 == 0 and stride == 1 and pixels >= 1024:
            return _strict_conv()
        return 0


_strict_conv_mode = None


def _strict_conv() -> int:
    """conv impl of the strict mode: 2 (three-term split, default), 3 (four-term), 0 (fp32 SIMT: VQB_STRICT_CONV=simt or no
    sm_100 device)"""
    global _strict_conv_mode
    if _strict_conv_mode is None:
        import os
        v = os.environ.get('VQB_STRICT_CONV', 'tc3').lower()
        mode = {'simt': 0, 'tc3': 2, 'tc4': 3}.get(v, 2)
        if mode and not lib.load().vqb_device_supports_tcgen05():
            mode = 0
        _strict_conv_mode = mode
    return _strict_conv_mode


def set_strict_conv(mode: str) -> None:
    """'tc3' | 'tc4' | 'simt' (A/B measurements and tests)"""
    global _strict_conv_mode
    _strict_conv_mode = {'simt': 0, 'tc3': 2, 'tc4': 3}[mode]


_precision = Precision('strict')
_weights_epoch = 0          # bumped whenever a kernel updates parameters behind autograd's back (optimizer step)


def set_precision(name: str) -> None:
    if name not in ('strict', 'fast'):
        raise ValueError(f'unknown precision mode {name!r}')
    _precision.name = name


def get_precision() -> Precision:
    return _precision


def bump_weights_epoch() -> None:
    global _weights_epoch
    _weights_epoch += 1


class ZeroArena:
    """Per-step pool of zero-initialised scratch (GroupNorm sum buffers, bias-gradient accumulators ...): the ~150 small
    torch.zeros fills of a training step become ONE memset at the start of the step (Trainer._step_body brackets the step with
    `zero_arena.step()`); outside a step, or when the pool is exhausted, zeros() is plain torch.zeros.  Buffers handed out are
    valid until the next step begins -- every user consumes them inside the forward / backward pass that took them."""

    def __init__(self, nbytes: int = 64 << 20):
        self.nbytes, self.buf, self.off, self.high, self.active = nbytes, None, 0, 0, False

    @contextlib.contextmanager
    def step(self, device):
        if self.active:                               # nested steps (a trainer driving another): the outer one owns the pool
            yield
            return
        if self.buf is None or self.buf.device != device:
            self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
            self.high = 0
        elif self.high:
            # everything ANY earlier step dirtied (monotone high-water mark: replayed CUDA graphs of other step variants write
            # the pool without passing through here)
            self.buf[:self.high].zero_()
        self.off, self.active = 0, True
        try:
            yield
        finally:
            self.active = False

    def zeros(self, n: int, dtype: torch.dtype, device) -> torch.Tensor:
        nb = n * torch.empty((), dtype=dtype).element_size()
        if not self.active or self.buf is None or self.buf.device != device or self.off + nb > self.nbytes:
            return torch.zeros(n, dtype=dtype, device=device)
        t = self.buf[self.off:self.off + nb].view(dtype)
        self.off = (self.off + nb + 255) & ~255
        self.high = max(self.high, min(self.off, self.nbytes))
        return t


zero_arena = ZeroArena()


class StepScalars:
    """A few fp32 scalars that change from step to step (learning rate, Adam bias corrections, Gumbel temperature / KL weight) and
    reach the kernels through DEVICE memory, so that a captured CUDA graph of the step reads this step's values.  upload() copies
    them host -> device on the current stream -- ordered before the step's kernels or its graph replay -- through a small ring of
    pinned slots, each guarded by an event: the host may run several steps ahead of the device without re-writing a slot whose
    queued copy has not executed yet (ONE pinned buffer read by a copy at execution time would hand a step the values of a later
    step whenever the host runs ahead, i.e. whenever the caller does not synchronise every step)."""
    SLOTS = 4

    def __init__(self, shape, device):
        self.dev = torch.zeros(shape, dtype=torch.float32, device=device)
        self._slots = [torch.zeros(shape, dtype=torch.float32).pin_memory() for _ in range(self.SLOTS)]
        self._done = [None] * self.SLOTS
        self._n = 0

    def upload(self, values: torch.Tensor) -> None:
        """values: CPU tensor of the buffer's shape.  A no-op under stream capture: the caller uploads before every replay."""
        if torch.cuda.is_current_stream_capturing():
            return
        s = self._n % self.SLOTS
        self._n += 1
        if self._done[s] is None:
            self._done[s] = torch.cuda.Event()
        else:
            self._done[s].synchronize()                    # the copy issued SLOTS uploads ago has read this slot
        self._slots[s].copy_(values)
        with torch.cuda.device(self.dev.device):
            self.dev.copy_(self._slots[s], non_blocking=True)
            self._done[s].record()


coop_cta_limit = 0          # cap on the grid of cooperative kernels (vqb_gn_bwd_fused); the data-parallel Trainer sets it to leave SMs to NCCL


_sink_off = 0


@contextlib.contextmanager
def no_grad_sink():
    """Around torch.autograd.grad(...) calls that ask for PARAMETER gradients (the adaptive generator weight, loss.py:80-96):
    inside, backward kernels return their parameter gradients to autograd instead of accumulating them into .grad."""
    global _sink_off
    _sink_off += 1
    try:
        yield
    finally:
        _sink_off -= 1


def grad_sink(p) -> Optional[torch.Tensor]:
    """The tensor a backward kernel may accumulate parameter `p`'s gradient into DIRECTLY (its .grad view inside FusedAdamW's
    flat gradient buffer) instead of returning it to autograd (whose AccumulateGrad is one `add` launch per parameter per step).
    Only when the optimizer opted the parameter in (`_vqb_direct_grad`, set by the single-process Trainer: the data-parallel
    bucket hooks are post-accumulate hooks and need autograd's accumulation) and no graph of the backward pass is recorded."""
    if p is None or _sink_off or not getattr(p, '_vqb_direct_grad', False) or torch.is_grad_enabled():
        return None
    g = p.grad
    return g if (g is not None and g.is_contiguous() and g.dtype == torch.float32) else None


def empty_nhwc(n: int, c: int, h: int, w: int, dtype: torch.dtype, device) -> torch.Tensor:
    return torch.empty((n, c, h, w), dtype=dtype, device=device, memory_format=CL)


def as_nhwc(t: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Return `t` as a channels-last tensor of `dtype` (no copy when it already is one)."""
    if not t.is_contiguous(memory_format=CL):
        t = t.contiguous(memory_format=CL)          # layout plumbing only (grad tensors produced outside our ops)
    if dtype is not None and t.dtype != dtype:
        out = torch.empty_like(t, dtype=dtype, memory_format=torch.preserve_format)
        call('vqb_convert', ptr(t), dt(t), ptr(out), dt(out), t.numel(), stream())
        t = out
    return t


def images_to_nhwc(images: torch.Tensor, dtype: torch.dtype, normalize: bool = True) -> torch.Tensor:
    """NCHW fp32 images -> channels-last; with `normalize` applies clamp[0,1] and (x-0.5)/0.5
    (reference: BaseVQVAE.preprocess_batch, abstract_modules/base_autoencoder.py:41-50, augmentation excluded)."""
    if images.dtype != torch.float32:
        images = images.float()
    images = images.contiguous()
    n, c, h, w = images.shape
    out = empty_nhwc(n, c, h, w, dtype, images.device)
    if normalize:
        call('vqb_nchw_to_nhwc', ptr(images), ptr(out), dt(out), n, c, h, w, 1, 0.0, 1.0, 0.5, 2.0, stream())
    else:
        call('vqb_nchw_to_nhwc', ptr(images), ptr(out), dt(out), n, c, h, w, 0, 0.0, 0.0, 0.0, 1.0, stream())
    return out


def nhwc_to_images(x: torch.Tensor, scale: float = 1.0, shift: float = 0.0, clamp: Optional[Tuple[float, float]] = None):
    """channels-last (any dtype) -> NCHW-contiguous fp32, y = x*scale+shift (+clamp)
    (reference: preprocess_visualization, base_autoencoder.py:52-61)."""
    x = as_nhwc(x)
    n, c, h, w = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    lo, hi = clamp if clamp is not None else (0.0, 0.0)
    call('vqb_nhwc_to_nchw', ptr(x), dt(x), ptr(out), n, c, h, w, scale, shift, int(clamp is not None), lo, hi, stream())
    return out


# ------------------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------------------
class _PackRegistry:
    """Every (weight, layout) pair the kernels asked for, with a PERSISTENT packed buffer.  When the weights epoch moves (an
    optimizer step rewrote the parameters), the first request re-packs ALL registered trainable entries in ONE launch
    (vqb_pack_conv_weights_batched) instead of ~100 single launches spread over the step; weights marked `_vqb_static` (frozen
    trunks: LPIPS) are packed once.  Entries hold weak references; a table is rebuilt when an entry is added or has died."""

    def __init__(self):
        self.entries = {}            # (id(weight), mode, dtype, scale) -> dict(ref, wp, shape, ...)
        self.table = None            # device uint8 tensor of PackDesc records
        self.table_keys = None
        self.total = 0
        self.epoch = -1              # weights epoch at which all table entries were last packed
        self.dirty = True
        self.table_host = None
        self.capture_refs = []       # what a stream capture baked into its graph by address (handed to the graph's owner)

    def _alive(self):
        return {k: e for k, e in self.entries.items() if e['ref']() is not None}

    def lookup(self, weight, mode, dtype, scale):
        import weakref
        k = (id(weight), mode, dtype, scale)
        e = self.entries.get(k)
        w = e['ref']() if e is not None else None
        if e is None or w is not weight or e['ptr'] != weight.data_ptr() or e['shape'] != tuple(weight.shape) or e['version'] != weight._version:
            co, ci, kh, kw = weight.shape
            numel = co * ci * kh * kw if mode < 4 else 64 * (co if mode == 4 else ci)      # modes 4/5 pad K to 64
            wp = e['wp'] if (e is not None and e['wp'].numel() == numel and e['wp'].device == weight.device) else \
                torch.empty(numel, dtype=dtype, device=weight.device)
            e = {'ref': weakref.ref(weight), 'wp': wp, 'ptr': weight.data_ptr(), 'shape': tuple(weight.shape), 'mode': mode, 'dtype': dtype,
                 'scale': scale, 'numel': numel, 'version': weight._version, 'fresh_epoch': None,
                 'static': bool(getattr(weight, '_vqb_static', False))}
            self.entries[k] = e
            if not e['static']:
                self.dirty = True
        return e

    def _pack_one(self, e, weight):
        co, ci, kh, kw = weight.shape
        src = weight.detach()
        if not src.is_contiguous() or src.dtype != torch.float32:
            src = src.float().contiguous()            # e.g. the transposed-codebook view used by the Gumbel einsum
        call('vqb_pack_conv_weight', ptr(src), ptr(e['wp']), dt(e['wp']), e['mode'], co, ci, kh, kw, e['scale'], stream())

    def _rebuild(self):
        import struct
        self.entries = self._alive()
        recs, start, keys = [], 0, []
        for k, e in self.entries.items():
            w = e['ref']()
            if e['static'] or not (w.is_contiguous() and w.dtype == torch.float32):
                continue                              # packed individually (once / through a temporary)
            co, ci, kh, kw = e['shape']
            recs.append(struct.pack('<QQiiiiiifiq', e['ptr'], e['wp'].data_ptr(), e['mode'], int(e['dtype'] == torch.bfloat16), co, ci, kh, kw,
                                    float(e['scale']), 0, start))
            keys.append(k)
            start += (e['numel'] + 4095) // 4096 * 4096          # one CTA serves 4096 elements of ONE record
        assert not recs or len(recs[0]) == lib.load().vqb_pack_desc_bytes()
        self.table_keys, self.total = keys, start
        dev = next(iter(self.entries.values()))['wp'].device if self.entries else None
        if recs:
            # pinned staging + async copy: legal under CUDA-graph capture (the copy becomes a node that re-reads `host` at replay)
            host = torch.frombuffer(bytearray(b''.join(recs)), dtype=torch.uint8).pin_memory()
            self.table = torch.empty(host.numel(), dtype=torch.uint8, device=dev)
            self.table.copy_(host, non_blocking=True)
            self.table_host = host
            if torch.cuda.is_current_stream_capturing():
                self.capture_refs.append((host, self.table))
        else:
            self.table = self.table_host = None
        self.dirty = False

    def get(self, weight, mode, dtype, scale):
        if not (weight.is_contiguous() and weight.dtype == torch.float32):
            # a temporary view (the transposed codebook of the Gumbel einsum): packed on the spot, never registered
            co, ci, kh, kw = weight.shape
            numel = co * ci * kh * kw if mode < 4 else 64 * (co if mode == 4 else ci)
            wp = torch.empty(numel, dtype=dtype, device=weight.device)
            src = weight.detach().float().contiguous()
            call('vqb_pack_conv_weight', ptr(src), ptr(wp), dt(wp), mode, co, ci, kh, kw, scale, stream())
            return wp
        e = self.lookup(weight, mode, dtype, scale)
        if e['static']:
            if e['fresh_epoch'] is None:
                self._pack_one(e, weight); e['fresh_epoch'] = -1
            return e['wp']
        if e['fresh_epoch'] == _weights_epoch:
            return e['wp']
        if not batched_pack_enabled():
            self._pack_one(e, weight); e['fresh_epoch'] = _weights_epoch
            return e['wp']
        if self.epoch != _weights_epoch and not self.dirty and self.table is not None and \
                any(self.entries[k]['ref']() is None for k in self.table_keys):
            self.dirty = True                         # a registered weight died: its storage may be unmapped by now, never launch on it
        if self.dirty:
            self._rebuild()
        if self.table is not None and self.epoch != _weights_epoch:
            call('vqb_pack_conv_weights_batched', ptr(self.table), len(self.table_keys), self.total, stream())
            if torch.cuda.is_current_stream_capturing():
                # the graph replays this launch: table, sources and destinations must outlive it (take_capture_refs)
                ents = [self.entries[k] for k in self.table_keys]
                self.capture_refs.append((self.table, self.table_host, [e['wp'] for e in ents], [e['ref']() for e in ents]))
            self.epoch = _weights_epoch
            for k in self.table_keys:
                self.entries[k]['fresh_epoch'] = _weights_epoch
        if e['fresh_epoch'] != _weights_epoch:        # registered after the batched launch of this epoch
            self._pack_one(e, weight); e['fresh_epoch'] = _weights_epoch
        return e['wp']


_pack_registry = _PackRegistry()


def take_capture_refs() -> list:
    """Tensors whose addresses the stream capture that just ended baked into its graph through this module's caches; the
    owner of the CUDA graph keeps the returned list for as long as the graph may be replayed."""
    refs, _pack_registry.capture_refs = _pack_registry.capture_refs, []
    return refs


_batched_pack = None


def batched_pack_enabled() -> bool:
    """VQB_BATCHED_PACK=0: one pack launch per weight and layout, as in round 1 (A/B measurements)"""
    global _batched_pack
    if _batched_pack is None:
        import os
        _batched_pack = os.environ.get('VQB_BATCHED_PACK', '1') != '0'
    return _batched_pack


def _packed_weight(weight: torch.Tensor, mode: int, dtype: torch.dtype, scale: float = 1.0) -> torch.Tensor:
    """Kernel-layout copy of an nn.Conv2d weight (see vqb_pack_conv_weight for the modes), refreshed when the weights changed."""
    return _pack_registry.get(weight, mode, dtype, scale)


def split_hi_lo(x: torch.Tensor) -> torch.Tensor:
    """fp32 channels-last [N,C,H,W] -> bf16 channels-last [N,2C,H,W] = [hi | lo] per pixel (vqb_split_hi_lo)"""
    x = as_nhwc(x, torch.float32)
    n, c, h, w = x.shape
    out = empty_nhwc(n, 2 * c, h, w, torch.bfloat16, x.device)
    call('vqb_split_hi_lo', ptr(x), ptr(out), n * h * w, c, stream())
    return out


def _is_split(x: torch.Tensor, c: int) -> bool:
    return x.dtype == torch.bfloat16 and x.shape[1] == 2 * c


def _packed_split_weight(weight: torch.Tensor, terms: int, dgrad: bool, scale: float = 1.0) -> torch.Tensor:
    """K-major bf16 pack of the split weight for impl 2 / 3: [wh | wh | wl (| wl)] along the contraction axis (input channels for
    the forward convolution, output channels -- with flipped taps -- for dgrad).  Cached per weight version like _packed_weight."""
    key = (weight._version, _weights_epoch, weight.data_ptr(), scale)
    cache = getattr(weight, '_vqb_pack', None)
    if cache is None or cache.get('key') != key:
        cache = {'key': key}
        try:
            weight._vqb_pack = cache
        except Exception:
            pass
    tag = ('split', terms, dgrad)
    if tag not in cache:
        w = weight.detach().float()
        if scale != 1.0:
            w = w * scale                        # equalised-lr gain folded in BEFORE the split (the product must split the scaled value)
        wh = w.bfloat16().float()
        wl = w - wh
        blocks = [wh, wh, wl] + ([wl] if terms == 4 else [])
        co, ci, kh, kw = w.shape
        if dgrad:
            w2 = torch.cat(blocks, dim=0).contiguous()                 # [terms*co, ci, kh, kw]
            wp = torch.empty(w2.numel(), dtype=torch.bfloat16, device=w.device)
            call('vqb_pack_conv_weight', ptr(w2), ptr(wp), BF16, 3, terms * co, ci, kh, kw, 1.0, stream())
        else:
            w2 = torch.cat(blocks, dim=1).contiguous()                 # [co, terms*ci, kh, kw]
            wp = torch.empty(w2.numel(), dtype=torch.bfloat16, device=w.device)
            call('vqb_pack_conv_weight', ptr(w2), ptr(wp), BF16, 2, co, terms * ci, kh, kw, 1.0, stream())
        cache[tag] = wp
    return cache[tag]


def _conv_fwd_raw(impl: int, x: torch.Tensor, wp: torch.Tensor, bias, residual, out_dtype, ci, co, kh, kw, pad, stride, act,
                  alpha, gain, gn_sums: Optional[torch.Tensor] = None, gn_groups: int = 0) -> torch.Tensor:
    n, _, h, w = x.shape
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (w + 2 * pad - kw) // stride + 1
    y = empty_nhwc(n, co, oh, ow, out_dtype, x.device)
    if gn_sums is not None:
        call('vqb_conv2d_fwd_gn', impl, ptr(x), dt(x), ptr(wp), ptr(bias), ptr(residual), ptr(y), dt(y), n, h, w, ci, co, kh, kw,
             pad, stride, act, alpha, gain, ptr(gn_sums), gn_groups, stream())
    else:
        call('vqb_conv2d_fwd', impl, ptr(x), dt(x), ptr(wp), ptr(bias), ptr(residual), ptr(y), dt(y), n, h, w, ci, co, kh, kw,
             pad, stride, act, alpha, gain, stream())
    return y


_gn_fusion = None


def gn_fusion_enabled() -> bool:
    """VQB_GN_FUSE=0 switches the fused GroupNorm statistics of the convolution epilogues off (A/B measurements)"""
    global _gn_fusion
    if _gn_fusion is None:
        import os
        _gn_fusion = os.environ.get('VQB_GN_FUSE', '1') != '0'
    return _gn_fusion


_narrowout = None


def _narrowout_enabled() -> bool:
    """VQB_NARROW_OUT=0: the 128 -> 3 head on the N = 16 halo kernel (A/B measurements)"""
    global _narrowout
    if _narrowout is None:
        import os
        _narrowout = os.environ.get('VQB_NARROW_OUT', '1') != '0'
    return _narrowout


_narrowin = None


def _narrowin_enabled() -> bool:
    """VQB_NARROW_IN=0: 3-channel inputs through the im2col tensor + 1x1 convolution (A/B measurements)"""
    global _narrowin
    if _narrowin is None:
        import os
        _narrowin = os.environ.get('VQB_NARROW_IN', '1') != '0'
    return _narrowin


def _packed_narrowout_weight(weight: torch.Tensor, w_scale: float) -> torch.Tensor:
    """[co <= 3, ci, 3, 3] -> [32][ci] bf16, row (tap * co + c) = w[c][:, tap] (vqb_conv2d_fwd_narrowout); persistent buffer,
    refreshed once per weights epoch"""
    cache = getattr(weight, '_vqb_narrowout', None)
    if cache is not None and cache['epoch'] == _weights_epoch and cache['scale'] == w_scale and cache['ptr'] == weight.data_ptr():
        return cache['wp']
    co, ci = weight.shape[0], weight.shape[1]
    rows = (weight.detach().float() * w_scale).permute(2, 3, 0, 1).reshape(9 * co, ci)
    if cache is None or cache['wp'].shape != (32, ci) or cache['wp'].device != weight.device:
        cache = {'wp': torch.zeros(32, ci, dtype=torch.bfloat16, device=weight.device)}
        weight._vqb_narrowout = cache
    cache['wp'][:9 * co].copy_(rows)
    cache.update(epoch=_weights_epoch, scale=w_scale, ptr=weight.data_ptr())
    return cache['wp']


def _narrow_route(prec: Precision, ci: int, co: int, kh: int, kw: int, pad: int, stride: int, frozen: bool = False) -> Optional[str]:
    """fast mode: the RGB heads (3x3, pad 1) ride the tensor cores as a 64-channel 1x1 implicit GEMM over an im2col tensor
    ('in': narrow input, e.g. encoder.conv_in 3->128; 'out': narrow output, e.g. decoder.conv_out 128->3, whose dgrad and
    wgrad are narrow-INPUT problems on dy)."""
    if prec.name != 'fast' or (kh, kw, pad, stride) != (3, 3, 1, 1):
        return None
    if ci <= 7 and (co % 128 == 0 or (frozen and co % 64 == 0)):       # the tcgen05 wgrad needs 128-wide Co tiles
        return 'in'
    if co <= 7 and ci % 128 == 0:
        return 'out'
    return None


_im2col_scratch = {}


def _im2col64(x: torch.Tensor) -> torch.Tensor:
    """64-channel im2col tensor of a narrow (C <= 7) input.  The tensor is consumed by the kernel launched right after it on
    the same stream and never saved for backward, so ONE zero-initialised scratch buffer per shape is reused and only the
    first ceil(9C/8)*8 columns are rewritten (the zero padding up to 64 is 58 % of the bytes for RGB)."""
    n, c, h, w = x.shape
    key = (n, h, w, c, x.device)
    p = _im2col_scratch.get(key)
    if p is None:
        if len(_im2col_scratch) >= 4:
            _im2col_scratch.clear()
        p = empty_nhwc(n, 64, h, w, torch.bfloat16, x.device).zero_()
        _im2col_scratch[key] = p
    call('vqb_im2col3x3_narrow', ptr(x), dt(x), ptr(p), BF16, n, h, w, c, 0, stream())
    return p


class ActBwdFn(torch.autograd.Function):
    """dpre = dy * act'(pre) * gain with act' recovered from the saved OUTPUT y (bias_act.py:143-210 with its `yref`).  Linear in
    dy, and for the piecewise-linear activations of the discriminator (lrelu) act'' = 0, so the backward of this function is
    the function itself; for the smooth activations only first order is built."""

    @staticmethod
    def forward(ctx, dy, y, act, alpha, gain, out_dtype, db=None):
        """`db` (optional, zero-initialised fp32 [C]) receives the bias gradient sum_p dpre[p, c] as a side effect when the
        fused single-pass kernel applies (one dtype throughout, C a multiple of the 16-byte vector); returns dpre."""
        dy = as_nhwc(dy) if dy.dim() == 4 else dy.contiguous()
        dpre = torch.empty_like(dy, dtype=out_dtype, memory_format=torch.preserve_format)
        c = dy.shape[1] if dy.dim() == 4 else 0
        vw = 8 if dy.dtype == torch.bfloat16 else 4
        ctx.db_done = False
        if dy.dim() == 4 and y.dtype == dy.dtype == out_dtype and c % vw == 0 and c // vw <= 256:
            call('vqb_act_bwd_bias', ptr(y), ptr(dy), ptr(dpre), dt(dy), act, alpha, gain, dy.numel() // c, c, ptr(db), stream())
            ctx.db_done = db is not None
        else:
            call('vqb_act_bwd_from_output', ptr(y), dt(y), ptr(dy), dt(dy), ptr(dpre), dt(dpre), act, alpha, gain, dy.numel(), stream())
        ActBwdFn.last_db_done = ctx.db_done
        ctx.save_for_backward(y)
        ctx.cfg = (act, alpha, gain, dy.dtype)
        return dpre

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        act, alpha, gain, dy_dtype = ctx.cfg
        if act not in (ACT_LRELU, ACT_RELU):
            raise lib.VQBError('second-order backward is only built for piecewise-linear activations')
        return ActBwdFn.apply(g, y, act, alpha, gain, dy_dtype), None, None, None, None, None, None


class Conv2dFn(torch.autograd.Function):
    """y = act(conv2d(x, w) + b) * gain + residual   (reference: nn.Conv2d in vqvae/modules/autoencoder.py:55-61,
    102,114,133,153,170; fused epilogues replace the separate `x + h` of ResBlock.forward :77 and torch.tanh :180)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, pad, stride, act, alpha, gain, out_dtype, w_scale, gn_groups=0):
        """gn_groups > 0: the output feeds GroupNorm(gn_groups, co); when the kernel can, its epilogue also accumulates the
        per-(image, group) sums of the output, returned as a second (non-differentiable) output [N * gn_groups * 2] float64
        (an empty tensor otherwise) that GroupNormActFn consumes instead of running its statistics pass."""
        prec = get_precision()
        co, ci, kh, kw = weight.shape
        impl = prec.conv_impl(ci, co, stride, x.shape[0] * x.shape[2] * x.shape[3]) if (pad == kh // 2 and kh == kw) else 0
        in_dtype = x.dtype
        if impl == 1:
            cdt = torch.bfloat16
        elif impl in (2, 3):
            cdt = torch.float32
        else:
            cdt = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else prec.act_dtype
        x = as_nhwc(x, cdt)
        out_dtype = out_dtype or prec.act_dtype
        if residual is not None:
            if act != ACT_NONE:
                # the activation derivative is recovered from the saved OUTPUT, which a fused residual would contaminate
                raise lib.VQBError('conv2d: a fused residual cannot be combined with an activation epilogue')
            residual = as_nhwc(residual, out_dtype)
        b = bias.detach().reshape(-1).float().contiguous() if bias is not None else None
        route = _narrow_route(prec, ci, co, kh, kw, pad, stride, frozen=not weight.requires_grad)
        n_, _, h_, w_ = x.shape
        sums = None

        def gn_buffer(eff_impl, eci, ekh, ekw, epad):
            if gn_groups and act == ACT_NONE and gn_fusion_enabled() and lib.load().vqb_conv2d_fwd_gn_supported(
                    eff_impl, n_, h_, w_, eci, co, ekh, ekw, epad, stride, gn_groups):
                return zero_arena.zeros(n_ * gn_groups * 2, torch.float64, x.device)
            return None

        if (route == 'out' and impl == 1 and x.dtype == torch.bfloat16 and residual is None and ci <= 512 and co <= 3 and
                (kh, kw, pad, stride) == (3, 3, 1, 1) and _narrowout_enabled()):
            # the 3-channel image head: per-tap partial products on the tensor cores + shift-add epilogue (one small GEMM per tile
            # instead of 72 N = 16 UMMAs)
            wp = _packed_narrowout_weight(weight, w_scale)
            y = empty_nhwc(n_, co, h_, w_, out_dtype, x.device)
            call('vqb_conv2d_fwd_narrowout', ptr(x), ptr(wp), ptr(b), ptr(y), dt(y), n_, h_, w_, ci, co, act, alpha, gain, stream())
        elif route == 'in' and ci == 3 and co <= 512 and x.dtype in (torch.float32, torch.bfloat16) and _narrowin_enabled():
            # the A operand (27 -> 64 im2col columns) is built in shared memory inside the kernel: no im2col tensor in HBM
            wp = _packed_weight(weight, 4, torch.bfloat16, w_scale)                  # [co][64], K zero-padded
            y = empty_nhwc(n_, co, h_, w_, out_dtype, x.device)
            call('vqb_conv2d_fwd_narrowin', ptr(x), dt(x), ptr(wp), ptr(b), ptr(residual), ptr(y), dt(y), n_, h_, w_, ci, co, act, alpha,
                 gain, stream())
        elif route == 'in':
            wp = _packed_weight(weight, 4, torch.bfloat16, w_scale)                  # [co][64], K zero-padded
            sums = gn_buffer(1, 64, 1, 1, 0)
            y = _conv_fwd_raw(1, _im2col64(x), wp, b, residual, out_dtype, 64, co, 1, 1, 0, 1, act, alpha, gain, sums, gn_groups)
        elif impl in (2, 3):
            # strict mode on the tensor cores: [hi | lo] operand halves, 3 / 4 bf16 products per multiply in one fp32 accumulator
            x = split_hi_lo(x)                                                        # saved in this form for the weight gradient
            wp = _packed_split_weight(weight, impl + 1, False, w_scale)
            sums = gn_buffer(impl, ci, kh, kw, pad)
            y = _conv_fwd_raw(impl, x, wp, b, residual, out_dtype, ci, co, kh, kw, pad, stride, act, alpha, gain, sums, gn_groups)
        else:
            wp = _packed_weight(weight, 2 if impl == 1 else 0, torch.bfloat16 if impl == 1 else torch.float32, w_scale)
            sums = gn_buffer(impl, ci, kh, kw, pad) if impl == 1 else None
            y = _conv_fwd_raw(impl, x, wp, b, residual, out_dtype, ci, co, kh, kw, pad, stride, act, alpha, gain, sums, gn_groups)
        ctx.save_for_backward(x, weight, y if act != ACT_NONE else None)
        ctx.cfg = (impl, pad, stride, act, alpha, gain, w_scale, bias is not None, residual is not None,
                   residual.dtype if residual is not None else None, in_dtype)
        if gn_groups:
            if sums is None:
                sums = torch.empty(0, dtype=torch.float64, device=y.device)
            ctx.mark_non_differentiable(sums)
            return y, sums
        return y

    @staticmethod
    def backward(ctx, dy, _g_sums=None):
        x, weight, y = ctx.saved_tensors
        impl, pad, stride, act, alpha, gain, w_scale, has_bias, has_res, res_dtype, in_dtype = ctx.cfg
        prec = get_precision()
        co, ci, kh, kw = weight.shape
        n, _, h, w = x.shape
        gdt = prec.act_dtype                            # storage dtype of activation gradients
        dy = as_nhwc(dy)
        dres = None
        if has_res:
            dres = dy if dy.dtype == res_dtype else as_nhwc(dy, res_dtype)
        db_fused = None
        if act != ACT_NONE:
            want_db = has_bias and ctx.needs_input_grad[2] and not _no_weight_grad and not torch.is_grad_enabled()
            db_buf = zero_arena.zeros(co, torch.float32, x.device) if want_db else None
            dy = ActBwdFn.apply(dy, y.detach(), act, alpha, gain, gdt, db_buf)
            if want_db and ActBwdFn.last_db_done:
                db_fused = db_buf                        # bias gradient produced by the same pass
        elif gain != 1.0:
            raise lib.VQBError('gain != 1 requires an activation epilogue')
        _, _, oh, ow = dy.shape
        dx = dw = db = None
        route = _narrow_route(prec, ci, co, kh, kw, pad, stride, frozen=not weight.requires_grad)
        if route is not None and torch.is_grad_enabled():
            raise lib.VQBError('conv2d: the RGB-head im2col routes are not twice differentiable')
        if route == 'in' and ctx.needs_input_grad[1] and not ctx.needs_input_grad[0]:
            # dW[(tap,ci)][co] = im2col(x)^T dy : 1x1 tcgen05 wgrad with 64 (zero-padded) input channels
            dwp = torch.zeros(64 * co, dtype=torch.float32, device=x.device)
            dyw = as_nhwc(dy, torch.bfloat16)
            if ci == 3 and co % 128 == 0 and x.dtype in (torch.float32, torch.bfloat16) and _narrowin_enabled():
                call('vqb_conv2d_wgrad_narrow', ptr(x), dt(x), ptr(dyw), ptr(dwp), n, h, w, ci, co, stream())     # im2col operand built in the kernel
            else:
                call('vqb_conv2d_wgrad', 1, ptr(_im2col64(x)), BF16, ptr(dyw), BF16, ptr(dwp), n, h, w, 64, co, 1, 1, 0, 1, stream())
            dw = torch.empty(weight.shape, dtype=torch.float32, device=x.device)
            call('vqb_unpack_conv_wgrad', ptr(dwp), ptr(dw), co, ci, kh, kw, w_scale, stream())
            if has_bias and ctx.needs_input_grad[2]:
                db = db_fused if db_fused is not None else _colsum(dy, n * oh * ow, co)
            return None, dw, db, dres, None, None, None, None, None, None, None, None
        if route == 'out' and x.dtype == torch.bfloat16:
            pd = None
            if ctx.needs_input_grad[0]:
                wd = _packed_weight(weight, 5, torch.bfloat16, w_scale)               # [ci][64]: flipped taps, K zero-padded
                ddt = in_dtype if in_dtype in (torch.float32, torch.bfloat16) else gdt
                if co == 3 and ci <= 512 and dy.dtype in (torch.float32, torch.bfloat16) and _narrowin_enabled():
                    # the input gradient of a Ci -> 3 head is a 3 -> Ci narrow-input convolution of dy: operand built inside the kernel
                    dx = empty_nhwc(n, ci, h, w, ddt, dy.device)
                    call('vqb_conv2d_fwd_narrowin', ptr(dy), dt(dy), ptr(wd), None, None, ptr(dx), dt(dx), n, h, w, co, ci, ACT_NONE, 0.0, 1.0,
                         stream())
                else:
                    pd = _im2col64(dy)                                                # im2col of the 3-channel gradient
                    dx = _conv_fwd_raw(1, pd, wd, None, None, ddt, 64, ci, 1, 1, 0, 1, ACT_NONE, 0.0, 1.0)
            if ctx.needs_input_grad[1]:
                # R[(kh',kw',co)][ci] = im2col(dy)^T x ; dW[co][ci][kh][kw] = R[(2-kh, 2-kw, co)][ci]
                r = torch.zeros(64 * ci, dtype=torch.float32, device=x.device)
                if co == 3 and ci % 128 == 0 and dy.dtype in (torch.float32, torch.bfloat16) and _narrowin_enabled():
                    call('vqb_conv2d_wgrad_narrow', ptr(dy), dt(dy), ptr(x), ptr(r), n, h, w, co, ci, stream())   # im2col(dy) built in the kernel
                else:
                    if pd is None:
                        pd = _im2col64(dy)
                    call('vqb_conv2d_wgrad', 1, ptr(pd), BF16, ptr(x), BF16, ptr(r), n, h, w, 64, ci, 1, 1, 0, 1, stream())
                dw = r[:9 * co * ci].view(3, 3, co, ci).flip(0, 1).permute(2, 3, 0, 1).contiguous()      # 10 KB reorder
                if w_scale != 1.0:
                    dw = dw * w_scale
            if has_bias and ctx.needs_input_grad[2]:
                db = db_fused if db_fused is not None else _colsum(dy, n * oh * ow, co)
            return dx, dw, db, dres, None, None, None, None, None, None, None, None
        if torch.is_grad_enabled():
            # the graph of this backward pass is being recorded (autograd.grad(..., create_graph=True): the R1 penalty,
            # loss.py:98-112): produce dx through ConvDgradFn, itself differentiable in dy and in the weight
            if (ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2])) and not _no_weight_grad:
                raise lib.VQBError('conv2d: the weight / bias gradients are not twice differentiable; record the backward graph under '
                                   'ops.no_weight_gradients() (as the reference does with conv2d_gradfix.no_weight_gradients)')
            if ctx.needs_input_grad[0]:
                ddt = in_dtype if in_dtype in (torch.float32, torch.bfloat16) else gdt
                dx = ConvDgradFn.apply(dy, weight, h, w, pad, stride, w_scale, ddt)
            return dx, None, None, dres, None, None, None, None, None, None, None, None
        dy_ops = dy
        if impl in (2, 3) and dy.dtype == torch.float32 and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            dy_ops = split_hi_lo(dy)                     # one [hi | lo] split of dy serves the input AND the weight gradient
        if ctx.needs_input_grad[0]:
            dx = _dgrad_raw(dy_ops, weight, h, w, pad, stride, w_scale, in_dtype if in_dtype in (torch.float32, torch.bfloat16) else gdt)
        if ctx.needs_input_grad[1] and not _no_weight_grad:
            dw = _wgrad_raw(x, dy_ops, weight.shape, pad, stride, w_scale, weight)
        if has_bias and ctx.needs_input_grad[2] and not _no_weight_grad:
            db = db_fused if db_fused is not None else _colsum(dy, n * oh * ow, co)
        return dx, dw, db, dres, None, None, None, None, None, None, None, None


def _colsum(dy: torch.Tensor, rows: int, co: int) -> torch.Tensor:
    db = torch.zeros(co, dtype=torch.float32, device=dy.device)
    call('vqb_colsum', ptr(dy), dt(dy), ptr(db), rows, co, stream())
    return db


def _dgrad_raw(dy: torch.Tensor, weight: torch.Tensor, h: int, w: int, pad: int, stride: int, w_scale: float, out_dtype):
    """dx [n, ci, h, w] of y = conv2d(x, weight * w_scale, pad, stride) given dy (channels-last)."""
    prec = get_precision()
    co, ci, kh, kw = weight.shape
    n = dy.shape[0]
    if stride != 1:
        # strided forward conv: transposed-conv gather over the (virtually) zero-upsampled dy, fp32 SIMT
        wd = _packed_weight(weight, 1, torch.float32, w_scale)
        dx = empty_nhwc(n, ci, h, w, out_dtype, dy.device)
        call('vqb_conv2d_dgrad', ptr(dy), dt(dy), ptr(wd), ptr(dx), dt(dx), n, h, w, ci, co, kh, kw, pad, stride, stream())
        return dx
    dimpl = prec.conv_impl(co, ci, 1, n * h * w) if kh - 1 - pad == kh // 2 else 0       # tcgen05 path needs a 'same' dgrad
    if _is_split(dy, co) and dimpl not in (2, 3):
        raise lib.VQBError('conv2d dgrad: split-precision gradient but no tensor-core kernel for this shape')
    if dimpl in (2, 3):
        dys = dy if _is_split(dy, co) else split_hi_lo(dy)
        wd = _packed_split_weight(weight, dimpl + 1, True, w_scale)
        return _conv_fwd_raw(dimpl, dys, wd, None, None, torch.float32 if out_dtype is None else out_dtype, co, ci, kh, kw,
                             kh - 1 - pad, 1, ACT_NONE, 0.0, 1.0)
    dyd = as_nhwc(dy, torch.bfloat16) if dimpl == 1 else dy
    wd = _packed_weight(weight, 3 if dimpl == 1 else 1, torch.bfloat16 if dimpl == 1 else torch.float32, w_scale)
    # dgrad = correlation of dy with the tap-flipped, channel-swapped weight; padding k-1-pad
    ddt = out_dtype if (dimpl == 0 or out_dtype == torch.float32) else prec.act_dtype
    return _conv_fwd_raw(dimpl, dyd, wd, None, None, ddt, co, ci, kh, kw, kh - 1 - pad, 1, ACT_NONE, 0.0, 1.0)


def _persistent_dwp(weight, numel: int) -> torch.Tensor:
    """The packed partial-sum buffer of `weight`'s gradient kernel, kept on the parameter: zero-filled ONCE -- the unpack kernel
    clears it behind its read (vqb_unpack_conv_wgrad_acc rezero), so no fill launch precedes a weight-gradient launch."""
    buf = getattr(weight, '_vqb_dwp', None)
    if buf is None or buf.numel() != numel or buf.device != weight.device or getattr(weight, '_vqb_dwp_dirty', False):
        buf = torch.zeros(numel, dtype=torch.float32, device=weight.device)
        weight._vqb_dwp = buf
    weight._vqb_dwp_dirty = True                       # cleared by the caller once the unpack (which re-zeroes) has been enqueued
    return buf


def _wgrad_raw(x: torch.Tensor, dy: torch.Tensor, wshape, pad: int, stride: int, w_scale: float, weight=None) -> Optional[torch.Tensor]:
    """dW [co, ci, kh, kw] fp32 of y = conv2d(x, W * w_scale, pad, stride) given x and dy (channels-last).  With `weight` (the
    parameter) the partial sums go through its persistent packed buffer, and -- when grad_sink(weight) allows -- the result is
    accumulated straight into weight.grad and None is returned."""
    prec = get_precision()
    co, ci, kh, kw = wshape
    n, _, h, w = x.shape
    simpl = prec.wgrad_impl(ci, co, stride, n * h * w)
    if simpl in (2, 3) and pad == kh // 2 and kh == kw and (_is_split(x, ci) or x.dtype == torch.float32):
        # split-precision operands: the bf16 kernel on [xh | xl] x [dyh | dyl] gives the four partial gradients hh, hl, lh, ll as
        # the four (co, ci) blocks of a (2co) x (2ci) problem; their sum is dW to ~2^-17 relative
        xs = x if _is_split(x, ci) else split_hi_lo(x)
        dys = dy if _is_split(dy, co) else split_hi_lo(dy)
        dwp = torch.zeros(kh * kw * 4 * ci * co, dtype=torch.float32, device=x.device)
        call('vqb_conv2d_wgrad', 1, ptr(xs), BF16, ptr(dys), BF16, ptr(dwp), n, h, w, 2 * ci, 2 * co, kh, kw, pad, stride, stream())
        dw4 = torch.empty((2 * co, 2 * ci, kh, kw), dtype=torch.float32, device=x.device)
        call('vqb_unpack_conv_wgrad', ptr(dwp), ptr(dw4), 2 * co, 2 * ci, kh, kw, w_scale, stream())
        return (dw4[:co, :ci] + dw4[:co, ci:]) + (dw4[co:, :ci] + dw4[co:, ci:])
    if _is_split(x, ci) and x.shape[1] != ci:
        raise lib.VQBError('conv2d wgrad: split-precision input but no tensor-core weight-gradient kernel for this shape')
    wimpl = prec.wgrad_impl(ci, co, stride) if x.dtype == torch.bfloat16 else 0
    wimpl = 1 if wimpl == 1 else 0
    dyw = as_nhwc(dy, torch.bfloat16) if wimpl == 1 else dy
    persistent = weight is not None and isinstance(weight, torch.nn.Parameter)
    dwp = _persistent_dwp(weight, kh * kw * ci * co) if persistent else torch.zeros(kh * kw * ci * co, dtype=torch.float32, device=x.device)
    call('vqb_conv2d_wgrad', wimpl, ptr(x), dt(x), ptr(dyw), dt(dyw), ptr(dwp), n, h, w, ci, co, kh, kw, pad, stride, stream())
    sink = grad_sink(weight) if persistent else None
    if sink is not None and tuple(sink.shape) == tuple(wshape):
        call('vqb_unpack_conv_wgrad_acc', ptr(dwp), ptr(sink), co, ci, kh, kw, w_scale, 1, 1, stream())
        weight._vqb_dwp_dirty = False
        return None
    dw = torch.empty(tuple(wshape), dtype=torch.float32, device=x.device)         # contiguous even if the weight is a view
    call('vqb_unpack_conv_wgrad_acc', ptr(dwp), ptr(dw), co, ci, kh, kw, w_scale, 0, int(persistent), stream())
    if persistent:
        weight._vqb_dwp_dirty = False
    return dw


_no_weight_grad = False


@contextlib.contextmanager
def no_weight_gradients():
    """conv2d_gradfix.no_weight_gradients (stylegan2_discriminator/ops/conv2d_gradfix.py): inside, conv backward passes
    produce input gradients only (used while the R1 penalty records the graph of a backward pass)."""
    global _no_weight_grad
    old, _no_weight_grad = _no_weight_grad, True
    try:
        yield
    finally:
        _no_weight_grad = old


class ConvDgradFn(torch.autograd.Function):
    """dx = d conv2d(x, w * w_scale) / dx applied to dy: bilinear in (dy, w).  Only instantiated while the graph of a backward
    pass is recorded; its own backward is a forward convolution (w.r.t. dy) and a weight-gradient kernel with the incoming
    second-order signal in the role of the layer input (w.r.t. w)."""

    @staticmethod
    def forward(ctx, dy, weight, h, w, pad, stride, w_scale, out_dtype):
        dy = as_nhwc(dy)
        ctx.save_for_backward(dy, weight)
        ctx.cfg = (pad, stride, w_scale)
        return _dgrad_raw(dy, weight, h, w, pad, stride, w_scale, out_dtype)

    @staticmethod
    def backward(ctx, ddx):
        dy, weight = ctx.saved_tensors
        pad, stride, w_scale = ctx.cfg
        g_dy = g_w = None
        cdt = get_precision().act_dtype
        if ctx.needs_input_grad[0]:
            g_dy = Conv2dFn.apply(ddx, weight, None, None, pad, stride, ACT_NONE, 0.0, 1.0, dy.dtype, w_scale)
        if ctx.needs_input_grad[1] and not _no_weight_grad:
            g_w = _wgrad_raw(as_nhwc(ddx.detach(), cdt), dy.detach(), weight.shape, pad, stride, w_scale)
        return g_dy, g_w, None, None, None, None, None, None


def conv2d(x, weight, bias=None, residual=None, pad=0, stride=1, act=ACT_NONE, alpha=0.0, gain=1.0, out_dtype=None,
           w_scale=1.0, gn_groups=0):
    """gn_groups > 0 (the output feeds GroupNorm(gn_groups, co)): the returned tensor carries `_gn_sums` = (sums, groups) when the
    convolution's epilogue produced the statistics, which group_norm_act() then uses instead of its own pass over the tensor."""
    if not gn_groups:
        return Conv2dFn.apply(x, weight, bias, residual, pad, stride, act, alpha, gain, out_dtype, w_scale)
    y, sums = Conv2dFn.apply(x, weight, bias, residual, pad, stride, act, alpha, gain, out_dtype, w_scale, gn_groups)
    if sums.numel():
        y._gn_sums = (sums, gn_groups)
    return y


# ------------------------------------------------------------------------------------------------------
# GroupNorm (+SiLU)
# ------------------------------------------------------------------------------------------------------
class GroupNormActFn(torch.autograd.Function):
    """act(GroupNorm(x)) with the reference's unbiased variance (vqvae/modules/autoencoder.py:25-39).

    With `want_skip` the input is also returned (as an identity output) for the ResBlock skip connection (`x + h`,
    autoencoder.py:77): the gradient arriving on that second output is then added inside the backward kernel instead of by
    a separate elementwise pass over the activation."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, act, want_skip, sums=None):
        x = as_nhwc(x)
        n, c, h, w = x.shape
        ga = gamma.detach().reshape(-1).float().contiguous()
        be = beta.detach().reshape(-1).float().contiguous()
        stats = torch.empty(n * groups * 2, dtype=torch.float32, device=x.device)
        if sums is None or sums.numel() != n * groups * 2:
            sums = zero_arena.zeros(n * groups * 2, torch.float64, x.device)
            call('vqb_gn_stats', ptr(x), dt(x), ptr(sums), n, h * w, c, groups, stream())
        # else: the producing convolution's epilogue already accumulated them (vqb_conv2d_fwd_gn)
        y = torch.empty_like(x, memory_format=torch.preserve_format)
        # mean / rstd are evaluated from the sums inside the apply kernel (and left in `stats` for the backward pass)
        call('vqb_gn_apply_sums', ptr(x), dt(x), ptr(sums), ptr(ga), ptr(be), ptr(y), dt(y), ptr(stats), n, h * w, c, groups, eps, act,
             stream())
        ctx.save_for_backward(x, stats, ga, be)
        ctx.params = (gamma, beta)
        ctx.cfg = (groups, act, gamma.shape, beta.shape)
        if want_skip:
            return y, x.view_as(x)
        return y

    @staticmethod
    def backward(ctx, dy, dskip=None):
        x, stats, ga, be = ctx.saved_tensors
        groups, act, gshape, bshape = ctx.cfg
        n, c, h, w = x.shape
        dy = as_nhwc(dy)
        gamma_p, beta_p = ctx.params
        if (ctx.needs_input_grad[0] and dy.dtype == x.dtype == torch.bfloat16 and
                lib.load().vqb_gn_bwd_fused_supported(BF16, BF16, BF16, n, h * w, c, groups, act)):
            # ONE cooperative launch: reduce image b, apply image b-1 out of L2 (x and dy cross HBM once instead of twice)
            buf = zero_arena.zeros(n * c * 2 + (n + 1) // 2, torch.float64, x.device)          # [part | per-image counters]
            part, counters = buf[:n * c * 2], buf[n * c * 2:].view(torch.int32)
            sg, sb = grad_sink(gamma_p), grad_sink(beta_p)
            direct = sg is not None and sb is not None and ctx.needs_input_grad[1] and ctx.needs_input_grad[2]
            if direct:
                dgamma, dbeta = sg.view(-1), sb.view(-1)
            else:
                dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
                dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
            dx = torch.empty_like(x, memory_format=torch.preserve_format)
            add = as_nhwc(dskip, dx.dtype) if dskip is not None else None
            call('vqb_gn_bwd_fused', ptr(x), ptr(dy), ptr(stats), ptr(ga), ptr(be), ptr(part), ptr(counters), ptr(add), ptr(dx),
                 ptr(dgamma), ptr(dbeta), int(direct), n, h * w, c, groups, act, coop_cta_limit, stream())
            if direct:
                return dx, None, None, None, None, None, None, None
            return dx, dgamma.reshape(gshape), dbeta.reshape(bshape), None, None, None, None, None
        part = zero_arena.zeros(n * c * 2, torch.float64, x.device)
        call('vqb_gn_bwd_reduce', ptr(x), dt(x), ptr(dy), dt(dy), ptr(stats), ptr(ga), ptr(be), ptr(part), n, h * w, c, groups,
             act, stream())
        if not ctx.needs_input_grad[0]:
            coef = torch.empty(n * groups * 2, dtype=torch.float32, device=x.device)
            dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
            dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
            call('vqb_gn_bwd_finalize', ptr(part), ptr(ga), ptr(coef), ptr(dgamma), ptr(dbeta), n, h * w, c, groups, stream())
            return None, dgamma.reshape(gshape), dbeta.reshape(bshape), None, None, None, None, None
        # the finalize arithmetic (per-group coefficients, dgamma / dbeta) runs inside the apply kernel; with a gradient sink the
        # parameter gradients are accumulated straight into the optimizer's flat buffer
        sg, sb = grad_sink(gamma_p), grad_sink(beta_p)
        direct = sg is not None and sb is not None and ctx.needs_input_grad[1] and ctx.needs_input_grad[2]
        if direct:
            dgamma, dbeta = sg.view(-1), sb.view(-1)
        else:
            dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
            dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x, memory_format=torch.preserve_format)
        add = as_nhwc(dskip, dx.dtype) if dskip is not None else None
        call('vqb_gn_bwd_apply_part', ptr(x), dt(x), ptr(dy), dt(dy), ptr(stats), ptr(ga), ptr(be), ptr(part), ptr(add), ptr(dx),
             dt(dx), ptr(dgamma), ptr(dbeta), int(direct), n, h * w, c, groups, act, stream())
        if direct:
            return dx, None, None, None, None, None, None, None
        return dx, dgamma.reshape(gshape), dbeta.reshape(bshape), None, None, None, None, None


def group_norm_act(x, gamma, beta, groups=32, eps=1e-6, act=ACT_SILU, want_skip=False):
    pre = getattr(x, '_gn_sums', None)
    sums = pre[0] if (pre is not None and pre[1] == groups) else None
    return GroupNormActFn.apply(x, gamma, beta, groups, eps, act, want_skip, sums)


# ------------------------------------------------------------------------------------------------------
# 2x resampling
# ------------------------------------------------------------------------------------------------------
def _down2(x, scale):
    n, c, h, w = x.shape
    y = empty_nhwc(n, c, h // 2, w // 2, x.dtype, x.device)
    call('vqb_down2', ptr(x), ptr(y), dt(x), n, h // 2, w // 2, c, scale, stream())
    return y


def _up2(x, scale):
    n, c, h, w = x.shape
    y = empty_nhwc(n, c, 2 * h, 2 * w, x.dtype, x.device)
    call('vqb_up2', ptr(x), ptr(y), dt(x), n, h, w, c, scale, stream())
    return y


class AvgPool2Fn(torch.autograd.Function):
    """F.avg_pool2d(x, 2, 2, 0)  (reference Downsample, autoencoder.py:89-91)."""

    @staticmethod
    def forward(ctx, x):
        x = as_nhwc(x)
        if x.shape[2] % 2 or x.shape[3] % 2:
            raise lib.VQBError('avg_pool2: odd spatial size')
        return _down2(x, 0.25)

    @staticmethod
    def backward(ctx, dy):
        return _up2(as_nhwc(dy), 0.25)


class Upsample2Fn(torch.autograd.Function):
    """F.interpolate(x, scale_factor=2, mode='nearest-exact')  (reference Upsample, autoencoder.py:103-105)."""

    @staticmethod
    def forward(ctx, x):
        return _up2(as_nhwc(x), 1.0)

    @staticmethod
    def backward(ctx, dy):
        return _down2(as_nhwc(dy), 1.0)


avg_pool2 = AvgPool2Fn.apply
upsample2 = Upsample2Fn.apply


# ------------------------------------------------------------------------------------------------------
# reconstruction losses
# ------------------------------------------------------------------------------------------------------
class DiffLossFn(torch.autograd.Function):
    """Returns (mean((a-b)^2), mean(|a-b|)) -- F.mse_loss (vqvae/model.py:274) and the L1/L2 terms of
    loss/loss.py:118-119.  `a_is_tanh` folds the tanh derivative of the producing conv epilogue into the gradient
    when `a` is marked as a tanh output (then the producer must NOT apply it again)."""

    @staticmethod
    def forward(ctx, a, b):
        a = as_nhwc(a)
        b = as_nhwc(b)
        sums = torch.zeros(2, dtype=torch.float64, device=a.device)
        call('vqb_diff_sums', ptr(a), dt(a), ptr(b), dt(b), ptr(sums), a.numel(), stream())
        ctx.save_for_backward(a, b)
        out = (sums / a.numel()).float()
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g2, g1):
        a, b = ctx.saved_tensors
        n = a.numel()
        # the upstream scalars stay on the device (no host sync): up = [g_l2, g_l1]
        up = torch.stack([g2.reshape(()), g1.reshape(())]).float().contiguous()
        da = torch.empty_like(a, memory_format=torch.preserve_format)
        call('vqb_diff_grad', ptr(a), dt(a), ptr(b), dt(b), ptr(da), dt(da), 1.0 / n, 1.0 / n, ptr(up), 0, n, stream())
        return da, None


def mse_l1(a, b):
    return DiffLossFn.apply(a, b)


@torch.no_grad()
def diff_sums(a, b) -> torch.Tensor:
    """-> fp64 [sum (a-b)^2, sum |a-b|] (the raw accumulators of vqb_diff_sums; evaluation metrics, model.py:491-562)"""
    a, b = as_nhwc(a), as_nhwc(b)
    sums = torch.zeros(2, dtype=torch.float64, device=a.device)
    call('vqb_diff_sums', ptr(a), dt(a), ptr(b), dt(b), ptr(sums), a.numel(), stream())
    return sums


def ssim_per_image(preds: torch.Tensor, target: torch.Tensor, k1: float = 0.01, k2: float = 0.03) -> torch.Tensor:
    """SSIM of every image of an NCHW fp32 batch, fp64 [N] (vqb_ssim_sums: torchmetrics' StructuralSimilarityIndexMeasure() with
    its defaults -- Gaussian 11 x 11 window, sigma 1.5, data_range=None = the larger of the two batch ranges; model.py:495, 529)."""
    preds, target = preds.float().contiguous(), target.float().contiguous()
    n, c, h, w = preds.shape
    if target.shape != preds.shape:
        raise lib.VQBError('ssim: shapes differ')
    (plo, phi), (tlo, thi) = torch.aminmax(preds), torch.aminmax(target)
    rng = torch.maximum(phi - plo, thi - tlo).reshape(1).float()
    sums = torch.zeros(n, dtype=torch.float64, device=preds.device)
    call('vqb_ssim_sums', ptr(preds), ptr(target), ptr(rng), ptr(sums), n, c, h, w, k1, k2, stream())
    return sums / float(c * (h - 10) * (w - 10))


# ------------------------------------------------------------------------------------------------------
# vector quantisation
# ------------------------------------------------------------------------------------------------------
class CodebookPrep:
    """fp16 copy and row norms of a codebook for the fused search (vqb_vq_fused), owned by a quantizer module and
    refreshed only when the codebook changed: the key holds the parameter's autograd version, its address and -- for codebooks
    an optimizer may rewrite behind autograd's back -- the global weights epoch.  The EMA update kernel refreshes the split
    itself (vq_ema_update), so the EMA path never runs a separate preparation launch."""

    def __init__(self):
        self.key = None
        self.hf = self.sq = None

    def _key(self, codebook: torch.Tensor):
        return (codebook._version, codebook.data_ptr(), tuple(codebook.shape), _weights_epoch if codebook.requires_grad else -1)

    def _alloc(self, codebook: torch.Tensor):
        k, d = codebook.shape
        if self.hf is None or self.hf.shape != (k, d) or self.hf.device != codebook.device:
            self.hf = torch.empty(k, d, dtype=torch.float16, device=codebook.device)
            self.sq = torch.empty(4 * k, dtype=torch.float32, device=codebook.device)     # norms + error-bound coefficients (vq_fused.cu)

    def get(self, codebook: torch.Tensor):
        key = self._key(codebook)
        if key != self.key:
            self._alloc(codebook)
            cb = codebook.detach()
            k, d = cb.shape
            call('vqb_vq_prep_codebook', ptr(cb), ptr(self.hf), ptr(self.sq), k, d, stream())
            self.key = key
        return self.hf, self.sq

    def mark_fresh(self, codebook: torch.Tensor):
        self.key = self._key(codebook)
        self._fresh_tag = True

    def invalidate_unless_fresh(self, codebook: torch.Tensor):
        """after a kernel rewrote the codebook behind autograd's back: keep the cached copy only if that kernel refreshed it too"""
        if not getattr(self, '_fresh_tag', False):
            self.key = None
        self._fresh_tag = False


def _fused_vq_ok(flat: torch.Tensor, k: int, d: int) -> bool:
    return d % 64 == 0 and d <= 256 and k % 8 == 0 and k <= 65528


def vq_stat_buffers(k: int, d: int, want_stats: bool, device):
    """ONE zero-filled fp32 buffer [counts (K) | dw (K*D)] (a single fill launch, and a single all-reduce for the EMA
    statistics under data parallelism) + the views into it."""
    buf = torch.zeros(k + (k * d if want_stats else 0), dtype=torch.float32, device=device)
    counts = buf[:k]
    dw = buf[k:].view(k, d) if want_stats else None
    return buf, counts, dw


def vq_assign_raw(flat: torch.Tensor, codebook: torch.Tensor, order: int, want_q: bool = True, want_stats: bool = False,
                  use_tc=None, prep: Optional[CodebookPrep] = None):
    """flat [N,D] fp32, codebook [K,D] fp32 -> (q or None, idx int64 [N], sse double[1], counts [K], dw [K,D] or None).
    use_tc: None = by precision mode (fast -> the one-launch fused kernel), True / 'fused' = fused kernel, 'legacy' = the round-1
    multi-launch tensor-core path (kept for A/B measurements), False = exact fp32 SIMT kernel (strict mode)."""
    n, d = flat.shape
    k = codebook.shape[0]
    dev = flat.device
    q = torch.empty_like(flat) if want_q else None
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    stat_buf, counts, dw = vq_stat_buffers(k, d, want_stats, dev)
    vq_assign_raw.last_stat_buffer = stat_buf
    if use_tc is None:
        use_tc = get_precision().name == 'fast'
    if use_tc and use_tc != 'legacy' and _fused_vq_ok(flat, k, d):
        scal = torch.zeros(2, dtype=torch.float64, device=dev)           # [sse | int32 x 2: re-ranked rows, full-scan rows]
        sse, und2 = scal[:1], scal[1:].view(torch.int32)
        und = und2[:1]
        vq_assign_raw.last_fullscan = und2[1:]
        hf, sq = (prep or CodebookPrep()).get(codebook)
        cb = codebook.detach()
        call('vqb_vq_fused', ptr(flat), ptr(cb), ptr(hf), ptr(sq), order, ptr(q), ptr(idx), ptr(sse), ptr(counts), ptr(dw),
             n, k, d, ptr(und), stream())
        vq_assign_raw.last_undecided = und
        return q, idx, sse, counts, dw
    sse = torch.zeros(1, dtype=torch.float64, device=dev)
    cb = codebook.detach()
    if use_tc and d % 64 == 0 and d <= 256 and k % 8 == 0 and k <= 8192:
        # round-1 path: tensor-core search + exact fp32 re-evaluation of near-tied rows over the whole code range (8 launches)
        ws_bytes = lib.load().vqb_vq_tc_workspace_bytes(n, k, d)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        und = torch.zeros(1, dtype=torch.int32, device=dev)
        call('vqb_vq_assign_tc', ptr(flat), ptr(cb), order, ptr(q), ptr(idx), ptr(sse), ptr(counts), ptr(dw), n, k, d, ptr(ws),
             ws_bytes, ptr(und), stream())
        vq_assign_raw.last_undecided = und
        return q, idx, sse, counts, dw
    ws_bytes = lib.load().vqb_vq_workspace_bytes(n, k, d)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    call('vqb_vq_assign', ptr(flat), ptr(cb), order, ptr(q), ptr(idx), ptr(sse), ptr(counts), ptr(dw), n, k, d, ptr(ws),
         ws_bytes, stream())
    return q, idx, sse, counts, dw


class VQFn(torch.autograd.Function):
    """Fused nearest-code quantisation with straight-through estimator and the MSE latent losses.

    forward(z [B,D,h,w] channels-last fp32, codebook [K,D]) ->
        (quantized [B,D,h,w], idx [B,h*w] int64, loss = (beta + cb_scale) * mean((e[idx]-z)^2),
         stats = ONE fp32 buffer [counts (K) | dw (K*D, only with want_stats)])
    reference: VectorQuantizer.forward vector_quantizers.py:23-61 (cb_scale=1), EMAVectorQuantizer.forward :128-180
    (cb_scale=0; the EMA state update itself is vqb_vq_ema_update, called by the module), EntropyVectorQuantizer :337-349.
    """

    @staticmethod
    def forward(ctx, z, codebook, order, beta, cb_scale, want_stats, prep=None):
        z = as_nhwc(z, torch.float32)
        b, d, h, w = z.shape
        flat = z.permute(0, 2, 3, 1).reshape(b * h * w, d)        # a view: channels-last memory is already (b h w) c
        if codebook.dtype != torch.float32 or not codebook.is_contiguous():
            raise lib.VQBError('the codebook must be a contiguous fp32 tensor')
        q, idx, sse, counts, dw = vq_assign_raw(flat, codebook, order, True, want_stats, prep=prep)
        stats = vq_assign_raw.last_stat_buffer
        loss = (sse[0] * ((beta + cb_scale) / flat.numel())).float()
        ctx.save_for_backward(flat, q, idx)
        ctx.cfg = (beta, cb_scale, codebook.shape, z.shape)
        qz = q.reshape(b, h, w, d).permute(0, 3, 1, 2)             # logical NCHW, physical NHWC
        ctx.mark_non_differentiable(idx, stats)
        return qz, idx.reshape(b, h * w), loss, stats

    @staticmethod
    def backward(ctx, g_q, _g_idx, g_loss, _g_stats):
        flat, q, idx = ctx.saved_tensors
        beta, cb_scale, cb_shape, zshape = ctx.cfg
        n, d = flat.shape
        k = cb_shape[0]
        gq = as_nhwc(g_q, torch.float32) if g_q is not None else None
        gl = g_loss.reshape(1).float().contiguous() if g_loss is not None else torch.zeros(1, device=flat.device)
        dz = torch.empty_like(flat)
        dcb = torch.zeros(cb_shape, dtype=torch.float32, device=flat.device) if (ctx.needs_input_grad[1] and cb_scale != 0.0) else None
        call('vqb_vq_backward', ptr(flat), ptr(q), ptr(idx), ptr(gq), ptr(gl), beta, cb_scale, ptr(dz), ptr(dcb), n, k, d,
             stream())
        b, _, h, w = zshape
        return dz.reshape(b, h, w, d).permute(0, 3, 1, 2), dcb, None, None, None, None, None


def vq_quantize(z, codebook, order=0, beta=0.25, cb_scale=1.0, want_stats=False, prep=None):
    return VQFn.apply(z, codebook, order, beta, cb_scale, want_stats, prep)


def code_histogram(idx: torch.Tensor, k: int) -> torch.Tensor:
    """int64 histogram of code indices WITHOUT a host synchronisation (torch.bincount reads the maximum back to size its
    output, which also forbids CUDA-graph capture): the per-batch usage counts of model.py:289-293."""
    out = torch.zeros(k, dtype=torch.int64, device=idx.device)
    flat = idx.reshape(-1)
    return out.scatter_add_(0, flat, torch.ones_like(flat))


def vq_codes(z, codebook, order=0, prep=None):
    """argmin only (vec_to_codes, vector_quantizers.py:63-84,182-203,358-381) -> idx [B, h*w] int64."""
    z = as_nhwc(z.detach(), torch.float32)
    b, d, h, w = z.shape
    flat = z.permute(0, 2, 3, 1).reshape(b * h * w, d)
    _, idx, _, _, _ = vq_assign_raw(flat, codebook, order, False, False, prep=prep)
    return idx.reshape(b, h * w)


def vq_ema_update(ema_count, ema_weight, codebook, counts, dw, decay, eps, batch, prep: Optional[CodebookPrep] = None):
    """EMA state update; with `prep` (fast mode) the same launch leaves the bf16 split / norms of the NEW codebook in it."""
    k, d = codebook.shape
    if prep is not None and get_precision().name == 'fast' and _fused_vq_ok(None, k, d):
        prep._alloc(codebook)
        call('vqb_vq_ema_update_prep', ptr(ema_count), ptr(ema_weight), ptr(codebook), ptr(counts), ptr(dw), ptr(prep.hf), ptr(prep.sq),
             k, d, decay, eps, float(batch), stream())
        prep.mark_fresh(codebook)
        return
    call('vqb_vq_ema_update', ptr(ema_count), ptr(ema_weight), ptr(codebook), ptr(counts), ptr(dw), k, d, decay, eps,
         float(batch), stream())


def vq_gather(codebook, idx):
    """codes_to_vec (abstract_modules/base_quantizer.py:53-61): idx [...,] int64 -> [..., D]."""
    cb = codebook.detach().float().contiguous()
    flat_idx = idx.reshape(-1).to(torch.int64).contiguous()
    out = torch.empty(flat_idx.numel(), cb.shape[1], dtype=torch.float32, device=cb.device)
    call('vqb_vq_gather', ptr(cb), ptr(flat_idx), ptr(out), flat_idx.numel(), cb.shape[0], cb.shape[1], stream())
    return out.reshape(*idx.shape, cb.shape[1])


def _gemm_nk(x2d: torch.Tensor, wp: torch.Tensor, k_in: int, k_out: int, residual=None, gain: float = 1.0) -> torch.Tensor:
    """[N,k_in] fp32 @ wp[k_in][k_out] fp32 (+ residual) * gain as a 1x1 'convolution' on N single-pixel images (fp32 SIMT
    implicit GEMM: the quantizers need fp32 distances / codebook gradients)."""
    n = x2d.shape[0]
    y = torch.empty(n, k_out, dtype=torch.float32, device=x2d.device)
    call('vqb_conv2d_fwd', 0, ptr(x2d), F32, ptr(wp), None, ptr(residual), ptr(y), F32, n, 1, 1, k_in, k_out, 1, 1, 0, 1, ACT_NONE,
         0.0, gain, stream())
    return y


def _entropy_tc_ok(codebook: torch.Tensor, n: int, k: int, d: int) -> bool:
    """VQB_ENTROPY_TC=0: the entropy quantizer's GEMMs on the fp32 SIMT kernels in fast mode too (A/B measurements)"""
    import os
    return (get_precision().name == 'fast' and d % 64 == 0 and k % 128 == 0 and n >= 1024 and codebook.dtype == torch.float32 and
            codebook.is_contiguous() and os.environ.get('VQB_ENTROPY_TC', '1') != '0' and bool(lib.load().vqb_device_supports_tcgen05()))


def _codebook_as_conv_weight(codebook: torch.Tensor) -> torch.Tensor:
    """[K,D] codebook as the weight [K,D,1,1] of a 1x1 convolution over the latent grid.  ONE view object per codebook storage, kept
    on the parameter: the packed-weight caches are keyed by the weight object."""
    v = getattr(codebook, '_vqb_view4', None)
    if v is None or v.data_ptr() != codebook.data_ptr() or v.shape[:2] != codebook.shape:
        v = codebook.detach().view(codebook.shape[0], codebook.shape[1], 1, 1)
        codebook._vqb_view4 = v
    return v


class VQEntropyFn(torch.autograd.Function):
    """EntropyVectorQuantizer.forward (vector_quantizers.py:290-356): nearest code on d = (|z|^2 - 2 z.e) + |e|^2, straight-through
    output, (1+beta)-weighted MSE terms and the entropy regulariser ratio * (mean_i H(p_i) - H(mean_i p_i)), p = softmax(-d/T).
    The N x K matrix exists once (fp32) and is transformed in place: dot -> log p (forward), log p -> dLoss/dd (backward)."""

    @staticmethod
    def forward(ctx, z, codebook, beta, ratio, temperature, argmax_target=False):
        z = as_nhwc(z, torch.float32)
        b, d, h, w = z.shape
        flat = z.permute(0, 2, 3, 1).reshape(b * h * w, d)
        cb = codebook.detach().float().contiguous()
        n, k = flat.shape[0], cb.shape[0]
        dev = flat.device
        use_tc = _entropy_tc_ok(codebook, n, k, d)
        if use_tc:
            # fast mode: the three GEMMs of this quantizer ride the tcgen05 1x1 convolution kernels.  The dot products feed a
            # softmax at T = 0.01 (a 100x gain on every distance error), so the forward product runs on SPLIT-PRECISION operands
            # (bf16 hi / lo halves, three products per multiply in one fp32 accumulator, ~2^-16 relative -- the strict mode's
            # convolution arithmetic); the two gradient products run on plain bf16 operands like every other gradient of this mode.
            w4 = _codebook_as_conv_weight(codebook)
            m4 = _conv_fwd_raw(2, split_hi_lo(z), _packed_split_weight(w4, 3, False), None, None, torch.float32, d, k, 1, 1, 0, 1,
                               ACT_NONE, 0.0, 1.0)                                       # [B,K,h,w] channels-last = [N,K]
            m = m4.permute(0, 2, 3, 1).reshape(n, k)
        else:
            cbt = torch.empty(d * k, dtype=torch.float32, device=dev)                   # E^T as a packed 1x1 weight [D][K]
            call('vqb_pack_conv_weight', ptr(cb), ptr(cbt), F32, 0, k, d, 1, 1, 1.0, stream())
            m = _gemm_nk(flat, cbt, d, k)                                                # dot products [N,K]
        cb_sq = torch.empty(k, dtype=torch.float32, device=dev)
        call('vqb_row_sqnorm', ptr(cb), ptr(cb_sq), k, d, stream())
        idx = torch.empty(n, dtype=torch.int64, device=dev)
        ent_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        call('vqb_vq_entropy_rows', ptr(m), ptr(flat), ptr(cb_sq), temperature, ptr(idx), ptr(ent_sum), n, k, d, int(argmax_target),
             stream())
        if argmax_target:
            # straight-through one-hot targets (:311-315): the batch-mean distribution is the code histogram
            colsum_p = torch.bincount(idx, minlength=k).float()
        else:
            colsum_p = torch.zeros(k, dtype=torch.float32, device=dev)
            call('vqb_vq_colsum_exp', ptr(m), ptr(colsum_p), n, k, stream())
        ent = torch.empty(2, dtype=torch.float32, device=dev)
        call('vqb_vq_entropy_finalize', ptr(colsum_p), ptr(ent_sum), ratio, ptr(ent), n, k, stream())
        q = torch.empty_like(flat)
        call('vqb_vq_gather', ptr(cb), ptr(idx), ptr(q), n, k, d, stream())
        sums = torch.zeros(2, dtype=torch.float64, device=dev)
        call('vqb_diff_sums', ptr(q), F32, ptr(flat), F32, ptr(sums), flat.numel(), stream())
        loss = (sums[0] * ((1.0 + beta) / flat.numel())).float() + ent[0]
        ctx.save_for_backward(flat, q, idx, m, colsum_p, cb)
        ctx.cfg = (beta, ratio, temperature, z.shape, bool(argmax_target))
        ctx.w4 = w4 if use_tc else None
        qz = q.reshape(b, h, w, d).permute(0, 3, 1, 2)
        ctx.mark_non_differentiable(idx)
        return qz, idx.reshape(b, h * w), loss

    @staticmethod
    def backward(ctx, g_q, _g_idx, g_loss):
        flat, q, idx, m, colsum_p, cb = ctx.saved_tensors
        beta, ratio, temperature, zshape, argmax_target = ctx.cfg
        n, d = flat.shape
        k = cb.shape[0]
        dev = flat.device
        gq = as_nhwc(g_q, torch.float32) if g_q is not None else None
        gl = g_loss.reshape(1).float().contiguous() if g_loss is not None else torch.zeros(1, device=dev)
        dz = torch.empty_like(flat)
        dcb = torch.zeros(k, d, dtype=torch.float32, device=dev)
        call('vqb_vq_backward', ptr(flat), ptr(q), ptr(idx), ptr(gq), ptr(gl), beta, 1.0, ptr(dz), ptr(dcb), n, k, d, stream())
        g = m.clone()                                                                    # keep logp intact for a second backward call
        call('vqb_vq_entropy_bwd_rows', ptr(g), ptr(colsum_p), ptr(gl), ratio, temperature, n, k, ptr(idx) if argmax_target else None,
             stream())
        b, _, h, w = zshape
        if ctx.w4 is not None:
            gb = as_nhwc(g.view(b, h, w, k).permute(0, 3, 1, 2), torch.bfloat16)         # [B,K,h,w] channels-last
            ge = _dgrad_raw(gb, ctx.w4, h, w, 0, 1, 1.0, torch.float32)                  # G E  [B,D,h,w]
            dz.add_(ge.permute(0, 2, 3, 1).reshape(n, d), alpha=-2.0)
            zb = as_nhwc(flat.view(b, h, w, d).permute(0, 3, 1, 2), torch.bfloat16)
            gtz = _wgrad_raw(zb, gb, (k, d, 1, 1), 0, 1, 1.0).reshape(k * d)              # G^T Z  [K][D]
        else:
            dz = _gemm_nk(g, cb, k, d, residual=dz, gain=-2.0)                           # dz += -2 G E   (wp[(k)][d] = E itself)
            gtz = torch.zeros(k * d, dtype=torch.float32, device=dev)
            call('vqb_conv2d_wgrad', 0, ptr(g), F32, ptr(flat), F32, ptr(gtz), n, 1, 1, k, d, 1, 1, 0, 1, stream())   # G^T Z
        colsum_g = torch.zeros(k, dtype=torch.float32, device=dev)
        call('vqb_colsum', ptr(g), F32, ptr(colsum_g), n, k, stream())
        call('vqb_vq_entropy_combine_dcb', ptr(dcb), ptr(cb), ptr(colsum_g), ptr(gtz), k, d, stream())
        return dz.reshape(b, h, w, d).permute(0, 3, 1, 2), dcb, None, None, None, None


def vq_entropy(z, codebook, beta, ratio, temperature, argmax_target=False):
    return VQEntropyFn.apply(z, codebook, beta, ratio, temperature, argmax_target)


class GumbelRowsFn(torch.autograd.Function):
    """Gumbel-softmax over the channel dimension of a channels-last [B,K,h,w] logits tensor + KL-to-uniform term
    (F.gumbel_softmax + vector_quantizers.py:234-241).  Returns (y [B,K,h,w], idx [B,h,w] int64, kl_mean)."""

    @staticmethod
    def forward(ctx, logits, exp_noise, tau, hard):
        logits = as_nhwc(logits, torch.float32)
        b, k, h, w = logits.shape
        n = b * h * w
        noise = as_nhwc(exp_noise, torch.float32) if exp_noise is not None else None
        y = torch.empty_like(logits, memory_format=torch.preserve_format)
        idx = torch.empty(n, dtype=torch.int64, device=logits.device)
        kl = torch.zeros(1, dtype=torch.float64, device=logits.device)
        if torch.is_tensor(tau):                 # temperature in device memory (CUDA-graph replay under a schedule)
            call('vqb_gumbel_rows_fwd_dev', ptr(logits), ptr(noise), ptr(tau), int(hard), ptr(y), ptr(idx), ptr(kl), n, k, stream())
        else:
            call('vqb_gumbel_rows_fwd', ptr(logits), ptr(noise), tau, int(hard), ptr(y), ptr(idx), ptr(kl), n, k, stream())
        ctx.save_for_backward(logits, noise)
        ctx.cfg = (tau, n, k)
        ctx.mark_non_differentiable(idx)
        return y, idx.reshape(b, h, w), (kl[0] / n).float()

    @staticmethod
    def backward(ctx, dy, _g_idx, g_kl):
        logits, noise = ctx.saved_tensors
        tau, n, k = ctx.cfg
        dyc = as_nhwc(dy, torch.float32) if dy is not None else None
        gk = g_kl.reshape(1).float().contiguous() if g_kl is not None else None
        dl = torch.empty_like(logits, memory_format=torch.preserve_format)
        if torch.is_tensor(tau):
            call('vqb_gumbel_rows_bwd_dev', ptr(logits), ptr(noise), ptr(tau), ptr(dyc), ptr(gk), 1.0 / n, ptr(dl), n, k, stream())
        else:
            call('vqb_gumbel_rows_bwd', ptr(logits), ptr(noise), tau, ptr(dyc), ptr(gk), 1.0 / n, ptr(dl), n, k, stream())
        return dl, None, None, None


def gumbel_rows(logits, exp_noise, tau, hard):
    return GumbelRowsFn.apply(logits, exp_noise, tau, hard)


# ------------------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------------------
def adamw_flat_dev(p, g, m, v, hyper, beta1, beta2, eps, weight_decay, grad_scale=1.0):
    """AdamW over a flat range with {lr, bias corrections} in device memory (hyper: fp32 [>=3]); the caller bumps the weights epoch"""
    call('vqb_adamw_dev', ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), ptr(hyper), beta1, beta2, eps, weight_decay, grad_scale, stream())


def adamw_flat(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    call('vqb_adamw', ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step, grad_scale, stream())
    bump_weights_epoch()
