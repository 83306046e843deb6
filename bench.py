#!/usr/bin/env python
"""bench.py -- images/sec of one full VQ-VAE / VQGAN train step (forward + backward + AdamW [+ discriminator step]) at 256x256.

    python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N ...          # the reference's CPU arithmetic (oracle port) on host cores

Headline workload (BASELINE.json configs[1], `--config cfg2`): example_confs/ema_vqvae.yaml with codebook 1024, 256x256
synthetic RGB, batch 64 per GPU (weak scaling), random-init weights in the reference's construction order, augmentation
off.  The same JSON line carries one sub-record per other north_star configuration under "configs" (each measured the same
way: device-resident value, end-to-end e2e, conv roofline, loss check):
    cfg3  entropy_vqvae.yaml, K=1024, B=64/GPU
    cfg4  gumbel_vqgan.yaml (Gumbel K=1024 + LPIPS-VGG16 + StyleGAN2 discriminator ACTIVE, R1 every 16th step), B=32/GPU
    cfg5  EMA VQGAN: gumbel_vqgan.yaml's loss / autoencoder with the EMA quantizer at K=8192, B=32/GPU
and "strict" = cfg2 in the fp32-parity numeric mode.  `--config X` makes X the headline and skips the sub-records;
`--only-headline` skips them for cfg2.  At N=1 the line also carries `cpu_baseline` (the oracle port on the host cores) and
`torch_eager_gpu_baseline` (the same port as stock torch eager ops on this GPU: fp32 / TF32 convolutions and bf16 autocast; none
of this repository's kernels).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import gc
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

UNIT = 'images/s'
FALLBACK_PEAKS = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}

WORKLOADS = {
    'cfg2': dict(yaml='ema_vqvae', codebook=1024, batch=64, quantizer=None, gan=False,
                 metric='images/sec VQ-VAE train step 256x256 (ema_vqvae, K=1024)',
                 label='ema_vqvae {S}x{S} K={K} B={B}/GPU (BASELINE configs[1]), fwd+bwd+AdamW, augmentation off'),
    'cfg3': dict(yaml='entropy_vqvae', codebook=1024, batch=64, quantizer=None, gan=False,
                 metric='images/sec VQ-VAE train step 256x256 (entropy_vqvae, K=1024)',
                 label='entropy_vqvae {S}x{S} K={K} B={B}/GPU (BASELINE configs[2]), fwd+bwd+AdamW, augmentation off'),
    'cfg4': dict(yaml='gumbel_vqgan', codebook=1024, batch=32, quantizer=None, gan=True,
                 metric='images/sec VQGAN train step 256x256 (gumbel_vqgan: Gumbel K=1024 + LPIPS-VGG16 + StyleGAN2 D)',
                 label='gumbel_vqgan {S}x{S} K={K} B={B}/GPU (BASELINE configs[3]): AE fwd+bwd+AdamW with L1/L2/LPIPS/generator '
                       'loss, then D fwd x2 + bwd + AdamW, R1 every 16th step, discriminator active (start_epoch 0), augmentation off'),
    'cfg5': dict(yaml='gumbel_vqgan', codebook=8192, batch=32, quantizer='ema', gan=True,
                 metric='images/sec VQGAN train step 256x256 (EMA VQGAN K=8192 + LPIPS-VGG16 + StyleGAN2 D)',
                 label='EMA VQGAN {S}x{S} K={K} B={B}/GPU (BASELINE configs[4]): gumbel_vqgan.yaml loss heads with the EMA quantizer '
                       'of ema_vqvae.yaml, same step as cfg4'),
}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            out = dict(FALLBACK_PEAKS)
            out.update({k: float(v) for k, v in p.items() if isinstance(v, (int, float))})
            return out, 'measured'
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), 'fallback'


def measured_traffic():
    """DRAM bytes per launch of the dominant conv kernel, parsed from THIS round's `ncu --set full` capture by
    tools/ncu_traffic.py into profiles/r02_conv_traffic.json (a profiler cannot run inside the timed region)."""
    path = os.path.join(ROOT, 'profiles', 'r02_conv_traffic.json')
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return None
    return None


def model_confs(name, args, world):
    from vqvae_vqgan_pytorch_lightning_b200.common_utils import derive_confs, get_model_conf
    wl = WORKLOADS[name]
    conf = get_model_conf(os.path.join(ROOT, 'example_confs', wl['yaml'] + '.yaml'))
    if wl['quantizer'] == 'ema':
        ema = get_model_conf(os.path.join(ROOT, 'example_confs', 'ema_vqvae.yaml'))['quantizer']
        conf['quantizer'] = dict(ema, embedding_dim=conf['quantizer']['embedding_dim'])
    batch = args.batch or wl['batch']
    codebook = args.codebook or wl['codebook']
    image_size, ae_conf, q_conf, l_conf, t_conf, bs = derive_confs(
        conf, world, {'num_embeddings': codebook, 'image_size': args.image_size, 'cumulative_bs': batch * world})
    if wl['gan']:
        l_conf = dict(l_conf)
        l_conf['adversarial_params'] = dict(l_conf['adversarial_params'], start_epoch=0)      # D active
    return image_size, ae_conf, q_conf, l_conf, t_conf, bs, codebook


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i',
                                          str(self.index), '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def mark(self):
        """index of the next sample: the sampler is started BEFORE the warm-up (nvidia-smi takes ~1 s to come up and holds a
        driver lock meanwhile -- inside a timed region that showed up as one 1000 ms step) and only the samples taken between
        mark() at the start of the timed regions and stop() are reported"""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.rows = self.rows[getattr(self, 'first', 0):]
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


CONV_ENTRY_POINTS = ('vqb_conv2d_fwd', 'vqb_conv2d_fwd_gn', 'vqb_conv2d_wgrad', 'vqb_conv2d_fwd_sub', 'vqb_conv2d_wgrad_sub',
                     'vqb_conv2d_fwd_narrowin', 'vqb_conv2d_fwd_narrowout', 'vqb_conv2d_wgrad_narrow')


def conv_shape(name, a):
    """(kind, impl, N, H, W, Ci, Co, k, stride, residual, ALGORITHMIC flop) of one convolution entry-point call.  fwd_gn = fwd + fused
    GroupNorm statistics.  The *_sub calls carry the discriminator's stride-2 3x3 convolution as a 2x2-tap convolution over the 4C
    channels of the space-to-depth input (16C MACs per output issued): they are counted with the 9C MACs per output of the
    strided convolution they implement (conv2d_resample.py:119-122), not with what the kernel issues."""
    if name in ('vqb_conv2d_fwd', 'vqb_conv2d_fwd_gn'):
        n, h, w, ci, co, kh, kw, pad, stride = a[8:17]
        oh, ow = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
        return ('fwd/dgrad' + ('+gn' if name.endswith('_gn') else ''), a[0], n, h, w, ci, co, kh, stride, bool(a[5]),
                2.0 * n * oh * ow * co * ci * kh * kw)
    if name == 'vqb_conv2d_wgrad':
        n, h, w, ci, co, kh, kw, pad, stride = a[6:15]
        oh, ow = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
        return ('wgrad', a[0], n, h, w, ci, co, kh, stride, False, 2.0 * n * oh * ow * co * ci * kh * kw)
    if name == 'vqb_conv2d_fwd_sub':
        n, hx, wx, h, w, ci, co, t, off = a[6:15]
        if off == 0:      # forward: [N, H, W, Co] from the 4C-channel space-to-depth input
            return ('s2d fwd', 1, n, h, w, ci, co, 3, 2, bool(a[3]), 2.0 * n * h * w * co * (ci // 4) * 9)
        return ('s2d dgrad', 1, n, h, w, ci, co, 3, 2, False, 2.0 * n * (h - 1) * (w - 1) * ci * (co // 4) * 9)
    if name == 'vqb_conv2d_fwd_narrowin':        # 3 -> Co image head, im2col operand built in the kernel
        n, h, w, ci, co = a[7:12]
        return ('fwd narrow-in', 1, n, h, w, ci, co, 3, 1, bool(a[4]), 2.0 * n * h * w * co * ci * 9)
    if name == 'vqb_conv2d_fwd_narrowout':       # Ci -> 3 image head, per-tap partial products + shift-add
        n, h, w, ci, co = a[5:10]
        return ('fwd narrow-out', 1, n, h, w, ci, co, 3, 1, False, 2.0 * n * h * w * co * ci * 9)
    if name == 'vqb_conv2d_wgrad_narrow':        # weight gradient of a 3-channel-sided head, im2col operand built in the kernel
        n, h, w, cn, cw = a[4:9]
        return ('wgrad narrow', 1, n, h, w, cn, cw, 3, 1, False, 2.0 * n * h * w * cn * cw * 9)
    if name == 'vqb_conv2d_wgrad_sub':
        n, hx, wx, h, w, ci, co, t, off = a[3:12]
        return ('s2d wgrad', 1, n, h, w, ci, co, 3, 2, False, 2.0 * n * h * w * co * (ci // 4) * 9)
    raise KeyError(name)


def conv_flops(name, a):
    return conv_shape(name, a)[-1]


# ------------------------------------------------------------------------------------------------------
def vq_microbench(pkg, dev, n_lat, k, use_tc, pk, iters=15):
    """The fused VQ op in isolation on a workload's latent shape: z ~ N(0,1) [N,256]; codebook N(0,1)
    (tie-free, what a trained codebook looks like to the search) and U(+-1/K) (the reference's initial codebook: thousands of
    codes within 1e-5 of each other, so the exact fp32 path takes over).  L2 is flushed between launches.  Both roofs:
    algorithmic bytes = 4ND (z) + 4KD (codebook) + 4ND (q) + 8N (idx) + 8K + 12KD (EMA statistics) against the HBM peak, and
    the distance contraction 2NKD (ONE fp16 product per multiply; exactness comes from the re-rank of near-ties) against the
    16-bit tensor peak."""
    d = 256
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    nbytes = 4 * n_lat * d * 2 + 4 * k * d + 8 * n_lat + 8 * k + 12 * k * d
    flop = 2.0 * n_lat * k * d
    cases = []
    for init in ('normal', 'uniform'):
        torch.manual_seed(0)
        z = torch.randn(n_lat, d, device=dev)
        cb = torch.randn(k, d, device=dev) if init == 'normal' else torch.empty(k, d, device=dev).uniform_(-1 / k, 1 / k)
        for _ in range(3):
            pkg.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=use_tc)
        ts, tk = [], []
        for _ in range(iters):
            flush.zero_()
            # the kernel launch alone (CUDA events around the C-ABI call: lib.KernelTimer) and the whole op as the module issues it
            # (+ the zero fill of the [counts | dw] statistics buffer and the output allocations)
            pkg.lib.timer = pkg.lib.KernelTimer(['vqb_vq_fused', 'vqb_vq_assign_tc', 'vqb_vq_assign'])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); pkg.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=use_tc); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
            rec = pkg.lib.timer.records; pkg.lib.timer = None
            tk.append(sum(a.elapsed_time(b) for _, _, a, b in rec) * 1e3 if rec else ts[-1])
        us_op = sorted(ts)[len(ts) // 2]
        us = sorted(tk)[len(tk) // 2]
        und = int(pkg.ops.vq_assign_raw.last_undecided) if use_tc else None
        cases.append({'codebook': init, 'us_per_launch': us, 'us_per_op_with_zero_fills': us_op, 'gbs': nbytes / us / 1e3,
                      'frac': nbytes / us / 1e3 / pk['hbm_gbs'],
                      'tensor_tflops_algorithmic': flop / us / 1e6, 'tensor_frac_algorithmic': flop / us / 1e6 / pk['bf16_tflops'],
                      'rows_on_exact_path': und})
    c0 = cases[0]
    return {'bound': 'hbm', 'kernel': 'vq_fused_kernel (vqb_vq_fused: one launch, tcgen05 search + exact re-rank + gather / STE / SSE / EMA sums)' if use_tc else 'vq_assign (fp32 SIMT)', 'achieved': c0['gbs'], 'peak': pk['hbm_gbs'],
            'unit': 'GB/s', 'frac': c0['frac'], 'traffic': None, 'algorithmic_bytes': nbytes, 'algorithmic_flop': flop,
            'shape': {'N': n_lat, 'K': k, 'D': d}, 'cases': cases}


def cpu_reference_step(name: str, batch: int, image_size: int, codebook: int, steps: int, warmup: int, threads: int):
    """The reference's arithmetic on host cores: oracle port of on_train_batch_start + training_step + AdamW (fp32 torch CPU,
    oneDNN / MKL) for workload `name`.  Returns (images_per_sec, seconds_per_step)."""
    from oracle import gan_oracle as G
    from oracle import init_state as oinit
    from vqvae_vqgan_pytorch_lightning_b200.common_utils import derive_confs, get_model_conf
    torch.set_num_threads(threads)

    class A:
        pass
    a = A(); a.batch, a.codebook, a.image_size = batch, codebook, image_size
    image_size, ae_conf, q_conf, l_conf, t_conf, bs, codebook = model_confs(name, a, 1)
    qtype = q_conf['type']
    crit = None if l_conf is None else 'gan'
    sd = oinit.init_state(qtype, codebook, q_conf['embedding_dim'], ae_conf['channels'], ae_conf['num_res_blocks'],
                          tuple(ae_conf['channel_multipliers']), seed=1234, criterion=crit, image_size=image_size)
    sd = oinit.make_leaf(sd, qtype)
    def num(v):                                    # PyYAML reads '1e-5' (no dot) as a string; the model casts with float() too
        try:
            return float(v) if isinstance(v, str) else v
        except ValueError:
            return v
    cfg = {'num_res_blocks': ae_conf['num_res_blocks'], 'channel_multipliers': tuple(ae_conf['channel_multipliers']),
           'quantizer': dict({k: num(v) for k, v in (q_conf.get('params') or {}).items()}, type=qtype)}
    opts = G.configure_optimizers(sd, t_conf, gan=(crit == 'gan'))
    x = torch.rand(batch, 3, image_size, image_size)
    times = []
    for it in range(warmup + steps):
        noise = torch.empty(batch, codebook, image_size // 16, image_size // 16).exponential_() if qtype == 'gumbel' else None
        t0 = time.perf_counter()
        G.train_step(sd, opts, x, cfg, l_conf, t_conf, 0, it, warmup + steps, exp_noise=noise)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch / sec, sec


class _TorchAdamW:
    """torch.optim.AdamW (foreach) behind the three calls gan_oracle.train_step makes on its optimizers -- what the reference's
    configure_optimizers returns (vqvae/model.py:424-440); the oracle's own per-tensor AdamW would be launch-bound on a GPU."""

    def __init__(self, sd, oracle_opt):
        self.lr = oracle_opt.lr
        self.opt = torch.optim.AdamW([{'params': [sd[n] for n in names], 'weight_decay': wd} for names, wd in oracle_opt.groups],
                                     lr=oracle_opt.lr, betas=oracle_opt.betas, eps=oracle_opt.eps, foreach=True)

    def zero_grad(self):
        self.opt.zero_grad(set_to_none=True)

    def step(self):
        for g in self.opt.param_groups:
            g['lr'] = self.lr
        self.opt.step()


def torch_eager_gpu_step(name: str, batch: int, image_size: int, codebook: int, steps: int, warmup: int, dev, autocast: bool):
    """Secondary baseline of SURVEY.md 8(d): the reference's arithmetic as stock torch EAGER ops on THIS GPU -- the oracle port of
    the reference modules with its tensors on the device (cuDNN convolutions, ATen element-wise kernels, cuBLAS distances),
    torch.optim.AdamW over the reference's parameter groups, inputs resident on the device.  autocast=True wraps the step in
    torch.autocast(bfloat16), the counterpart of the reference's Trainer(precision='16-mixed') (vqvae/train.py:129).  None of this
    repository's kernels run here.  Returns (images_per_sec, ms_per_step, peak GiB)."""
    from oracle import gan_oracle as G
    from oracle import init_state as oinit

    class A:
        pass
    a = A(); a.batch, a.codebook, a.image_size = batch, codebook, image_size
    image_size, ae_conf, q_conf, l_conf, t_conf, bs, codebook = model_confs(name, a, 1)
    qtype = q_conf['type']
    crit = None if l_conf is None else 'gan'
    sd = oinit.init_state(qtype, codebook, q_conf['embedding_dim'], ae_conf['channels'], ae_conf['num_res_blocks'],
                          tuple(ae_conf['channel_multipliers']), seed=1234, criterion=crit, image_size=image_size)
    sd = oinit.make_leaf({k: v.to(dev) for k, v in sd.items()}, qtype)

    def num(v):
        try:
            return float(v) if isinstance(v, str) else v
        except ValueError:
            return v
    cfg = {'num_res_blocks': ae_conf['num_res_blocks'], 'channel_multipliers': tuple(ae_conf['channel_multipliers']),
           'quantizer': dict({k: num(v) for k, v in (q_conf.get('params') or {}).items()}, type=qtype)}
    opts = [_TorchAdamW(sd, o) for o in G.configure_optimizers(sd, t_conf, gan=(crit == 'gan'))]
    x = torch.rand(batch, 3, image_size, image_size, device=dev)
    torch.cuda.reset_peak_memory_stats(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(warmup + steps):
        if it == warmup:
            torch.cuda.synchronize(dev)
            ev[0].record()
        noise = (torch.empty(batch, codebook, image_size // 16, image_size // 16, device=dev).exponential_()
                 if qtype == 'gumbel' else None)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
            G.train_step(sd, opts, x, cfg, l_conf, t_conf, 0, it, warmup + steps, exp_noise=noise)
    ev[1].record()
    torch.cuda.synchronize(dev)
    ms = ev[0].elapsed_time(ev[1]) / steps
    return batch / ms * 1e3, ms, torch.cuda.max_memory_allocated(dev) / 2 ** 30


def torch_eager_gpu_baseline(name: str, image_size: int, codebook: int, dev, batch: int = 32, steps: int = 3, warmup: int = 2):
    import gc
    rec = {'unit': UNIT, 'kind': 'oracle port of the reference modules as stock torch eager ops on the same GPU (cuDNN / ATen / cuBLAS, '
                                  'torch.optim.AdamW foreach); none of this repository\'s kernels; inputs resident on the device',
           'micro_batch': batch, 'steps': steps, 'warmup': warmup}
    for tag, autocast in (('fp32_tf32_convs', False), ('bf16_autocast', True)):
        b = batch
        while True:
            gc.collect(); torch.cuda.empty_cache()
            try:
                ips, ms, gib = torch_eager_gpu_step(name, b, image_size, codebook, steps, warmup, dev, autocast)
                rec[tag] = {'value': ips, 'ms_per_step': ms, 'micro_batch': b, 'peak_mem_gib': gib}
                break
            except torch.OutOfMemoryError:
                if b <= 8:
                    rec[tag] = {'error': 'out of memory at micro-batch 8'}
                    break
                b //= 2
            except Exception as e:
                rec[tag] = {'error': f'{type(e).__name__}: {e}'[:300]}
                break
    gc.collect(); torch.cuda.empty_cache()
    return rec


def run_reference(args, rank, world, out=sys.stdout):
    if rank != 0:
        return
    name = args.config
    wl = WORKLOADS[name]
    threads = os.cpu_count() or 1
    b = args.cpu_batch
    codebook = args.codebook or wl['codebook']
    ips, sec = cpu_reference_step(name, b, args.image_size, codebook, args.steps, min(args.warmup, 1), threads)
    line = {
        'impl': 'reference', 'metric': wl['metric'], 'value': ips, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': min(args.warmup, 1), 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['label'].format(S=args.image_size, K=codebook, B=args.batch or wl['batch']) +
                               f' -- bounded sample at micro-batch {b} (img/s is batch-insensitive on CPU)'},
        'cpu_baseline': {'value': ips, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': f'{args.steps} steps x {b} images, oracle port of the reference modules (fp32, torch CPU)'},
        'e2e': {'value': ips, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    out.write(json.dumps(line) + '\n'); out.flush()


# ------------------------------------------------------------------------------------------------------
def run_workload(name, args, pkg, dev, rank, world, precision, steps, warmup, stamp, sample_clocks=False, ncu_step=False):
    """Build workload `name`, run `warmup` + `steps` device-resident steps and `steps` end-to-end steps through
    Trainer.run_step; returns the record (rank 0; other ranks return None).  With --graph (default) the trainer replays one
    captured CUDA graph per step; a third, EAGER timed region of the same length then times the convolution entry points with
    CUDA events (per-kernel events cannot be recorded inside a graph replay) for the roofline object."""
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    pkg.set_precision(precision)
    image_size, ae_conf, q_conf, l_conf, t_conf, bs, codebook = model_confs(name, args, world)
    torch.manual_seed(1234)                       # same weights on every rank (reference construction order)
    model = pkg.VQVAE(image_size, ae_conf, q_conf, l_conf, t_conf, pretrained_lpips=False).to(dev).train()
    use_graph = bool(args.graph) and not ncu_step
    trainer = Trainer(max_epochs=1, num_training_batches=steps * 3 + warmup + 4, cuda_graph=use_graph)
    trainer.attach(model)
    model.on_train_start()
    model.training_augmentations = None           # the metric is quoted with augmentation off

    torch.manual_seed(1234 + rank)
    nbuf = 2
    host = [torch.rand(bs, 3, image_size, image_size).pin_memory() for _ in range(nbuf)]
    resident = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_idx = [0]
    l2_log = []                                    # device scalars, read after the timed region

    def step_resident(i):
        loss = trainer.run_step(resident[i % nbuf], step_idx[0]); step_idx[0] += 1
        # (clones: under graph replay the step's outputs live in static tensors that the next replay overwrites)
        l2_log.append((loss.detach().clone(), torch.as_tensor(model.logged['train/l2_loss']).detach().clone()))
        return loss

    def step_e2e(x):
        """x: this step's batch, already requested from pinned host memory by the DevicePrefetcher (the copy of the NEXT batch
        overlaps this step, as a DataLoader(pin_memory=True) does for the reference); every step's H2D copy and D2H loss read
        happen inside the timed region"""
        loss = trainer.run_step(x, step_idx[0]); step_idx[0] += 1
        l2_log.append((loss.detach().clone(), torch.as_tensor(model.logged['train/l2_loss']).detach().clone()))
        return float(loss.detach().cpu())                                        # D2H read of the step's loss

    stamp(f'{name}/{precision}: model and inputs ready')
    clocks = ClockSampler(dev.index or 0)
    if rank == 0 and sample_clocks:
        clocks.start()
    # warm-up: W steps, plus -- with graphs -- enough steps for every step variant met in the timed regions to be captured
    # (2 eager steps + the capture step per variant; the R1 variant of the VQGAN configs comes every 16th step)
    for i in range(warmup + ((3 + (33 if WORKLOADS[name]['gan'] else 0)) if use_graph else 0)):
        step_resident(i)
    barrier()
    if ncu_step:
        torch.cuda.profiler.start()
        step_resident(0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None

    # ---- timed region 1: inputs resident in HBM (value, ms_per_step) ---------------------------------------------
    if rank == 0 and sample_clocks:
        clocks.mark()
    if not use_graph:
        pkg.lib.timer = pkg.lib.KernelTimer(list(CONV_ENTRY_POINTS) + ['vqb_vq_assign', 'vqb_vq_assign_tc', 'vqb_vq_fused'])
    launches0 = pkg.lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        step_resident(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pkg.lib.launch_count - launches0
    ksum = pkg.lib.timer.summary() if pkg.lib.timer is not None else {}
    pkg.lib.timer = None

    # ---- timed region 2: end to end through the public API with host buffers ---------------------------------------
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import DevicePrefetcher
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_e2e(next(DevicePrefetcher([host[0]], dev)))        # untimed: first use of the pinned H2D / D2H staging path
    feed = DevicePrefetcher((host[i % nbuf] for i in range(steps)), dev).preallocate(host[0])     # ring allocated; no copy issued yet
    barrier()
    f0.record()                                    # exactly `steps` H2D copies follow, all inside the timed region
    t_dbg = []
    for i in range(steps):
        t_a = time.perf_counter()
        step_e2e(next(feed))
        t_dbg.append((time.perf_counter() - t_a) * 1e3)
    f1.record()
    stamp(f'{name}/{precision}: e2e per-step wall ms ' + ' '.join(f'{v:.1f}' for v in t_dbg))
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    clk = clocks.stop() if (rank == 0 and sample_clocks) else None
    ms_eager = None
    if use_graph:
        # ---- timed region 3 (graphs only): the same steps EAGERLY, every convolution entry point bracketed by CUDA events ------
        trainer.cuda_graph = False
        step_resident(0)
        pkg.lib.timer = pkg.lib.KernelTimer(list(CONV_ENTRY_POINTS) + ['vqb_vq_assign', 'vqb_vq_assign_tc', 'vqb_vq_fused'])
        launches0 = pkg.lib.launch_count
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        h0.record()
        for i in range(steps):
            step_resident(i)
        h1.record()
        barrier()
        ms_eager = h0.elapsed_time(h1)
        launches = pkg.lib.launch_count - launches0
        ksum = pkg.lib.timer.summary()
        pkg.lib.timer = None
    stamp(f'{name}/{precision}: timed regions done')

    t = torch.tensor([ms, ms_e2e, ms_eager or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    ms_eager = float(t[2]) if ms_eager is not None else None

    # ---- the step must be a real optimisation step: every loss finite, reconstruction error not increasing ----------
    losses = torch.stack([a.float().reshape(()) for a, _ in l2_log]).cpu().tolist()
    l2s = torch.stack([b.float().reshape(()) for _, b in l2_log]).cpu().tolist()
    if not all(math.isfinite(v) for v in losses + l2s):
        raise SystemExit(f'bench.py: non-finite loss in {name}/{precision}: {losses}')
    head, tail = sum(l2s[:nbuf]) / nbuf, sum(l2s[-nbuf:]) / nbuf          # the same nbuf batches at the start and at the end
    if tail > head * 1.001:
        raise SystemExit(f'bench.py: reconstruction loss increased in {name}/{precision}: {head} -> {tail}')
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
    del model, trainer, resident, host
    gc.collect()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    if rank != 0:
        return None

    pk, pk_kind = peaks()
    total_images = bs * world * steps
    conv_ms = sum(ksum[k]['ms'] for k in CONV_ENTRY_POINTS if k in ksum)
    fast = precision != 'strict'
    # strict mode: the weight gradient of a split-precision layer is ONE bf16 launch on [hi | lo] operands with both channel counts
    # doubled (four partial products): count it with the FLOPs of the convolution it implements
    def algo(k, a):
        return conv_flops(k, a) / (4.0 if (not fast and k == 'vqb_conv2d_wgrad' and a[0] == 1) else 1.0)
    conv_fl = sum(algo(k, a) for k in CONV_ENTRY_POINTS if k in ksum for a in ksum[k]['args'])
    conv_calls = sum(ksum[k]['calls'] for k in CONV_ENTRY_POINTS if k in ksum)
    achieved = conv_fl / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    peak = pk['bf16_tflops_sustained']
    tr = measured_traffic() if (name == 'cfg2' and fast and (image_size, bs) == (256, 64)) else None
    roofline = {'bound': 'tensor', 'kernel': 'implicit-GEMM conv (fwd+dgrad+wgrad launches)' + ('' if fast else ' -- split-precision bf16 hi/lo operands through the same tcgen05 kernels (3 tensor-core passes per algorithmic FLOP, fp32 storage), ALGORITHMIC FLOPs against the bf16 roof'),
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'traffic': tr['dram_bytes_per_launch'] if tr else None, 'traffic_source': tr.get('source') if tr else None,
                'traffic_algorithmic_bytes': tr.get('algorithmic_bytes') if tr else None,
                'peak_source': f'{pk_kind} (sustained bf16)', 'launches_timed': conv_calls,
                'share_of_step': conv_ms / (ms_eager if ms_eager is not None else ms),
                'measured_in': (f'eager timed region of {steps} steps ({ms_eager / steps:.2f} ms/step): CUDA events around every conv entry '
                                f'point; the headline value replays a CUDA graph of the same step') if ms_eager is not None else
                               'the headline timed region (CUDA events around every conv entry point)',
                'flop_per_step': conv_fl / steps, 'whole_step_tflops': conv_fl / (ms / 1e3) / 1e12,
                'whole_step_frac': conv_fl / (ms / 1e3) / 1e12 / peak}
    rec = {
        'metric': WORKLOADS[name]['metric'], 'value': total_images / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms / steps,
        'steps': steps, 'warmup': warmup, 'dtype': 'bf16' if fast else 'f32',
        'config': {'workload': WORKLOADS[name]['label'].format(S=image_size, K=codebook, B=bs), 'precision': precision,
                   'global_batch': bs * world, 'parallelism': f'dp{world}',
                   'l2_policy': 'working set per step (inputs + activations > 5 GB) exceeds the 126 MB L2'},
        'e2e': {'value': total_images / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': bs * 3 * image_size * image_size * 4,
                'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / steps},
        'gpu_launches': launches, 'cuda_graph': use_graph, 'ms_per_step_eager': (ms_eager / steps) if ms_eager is not None else None,
        'roofline': roofline,
        'loss': {'first': losses[0], 'last': losses[-1], 'l2_first': head, 'l2_last': tail, 'all_finite': True,
                 'steps_observed': len(losses)},
        'peak_mem_gib': peak_mem,
    }
    vq_step = ksum.get('vqb_vq_fused') or ksum.get('vqb_vq_assign_tc') or ksum.get('vqb_vq_assign')
    if vq_step:
        rec['vq_in_step_us_per_launch'] = vq_step['ms'] / vq_step['calls'] * 1e3
        und, full = getattr(pkg.ops.vq_assign_raw, 'last_undecided', None), getattr(pkg.ops.vq_assign_raw, 'last_fullscan', None)
        if und is not None and full is not None and ksum.get('vqb_vq_fused'):
            # the first steps after initialisation are the search's worst case: the reference's U(+-1/K) codebook against
            # large-norm latents puts hundreds of codes within fp32 rounding of each other (rows that need the full exact scan)
            rec['vq_in_step_rows'] = {'re_ranked': int(und), 'full_scan': int(full), 'of': bs * (image_size // 16) ** 2}
    if clk is not None:
        rec['clocks'] = clk
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='cfg2', choices=list(WORKLOADS))
    ap.add_argument('--only-headline', action='store_true', help='skip the cfg3 / cfg4 / cfg5 / strict sub-records')
    ap.add_argument('--precision', default=os.environ.get('VQB_PRECISION', 'fast'), choices=['strict', 'fast'])
    ap.add_argument('--batch', type=int, default=0, help='images per GPU per step (default: the workload\'s)')
    ap.add_argument('--image-size', type=int, default=256)
    ap.add_argument('--codebook', type=int, default=0)
    ap.add_argument('--cpu-batch', type=int, default=2)
    ap.add_argument('--cpu-steps', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-eager-baseline', action='store_true',
                    help='skip the secondary baseline (the oracle port of the reference modules as torch eager ops on the same GPU)')
    ap.add_argument('--graph', type=int, default=1, help='1 (default): replay one captured CUDA graph per step; 0: eager launches')
    ap.add_argument('--ncu-step', action='store_true', help='after the warm-up run ONE step between cudaProfilerStart/Stop and exit '
                    '(for `ncu --profile-from-start off`; prints no bench line)')
    args = ap.parse_args()

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    t_start = time.time()
    # stdout carries exactly ONE JSON line: everything libraries print to fd 1 (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)

    def stamp(what):                              # progress on stderr (the JSON line is the only thing on stdout)
        if os.environ.get('VQB_BENCH_VERBOSE'):
            print(f'[bench rank {rank} +{time.time() - t_start:6.1f}s] {what}', file=sys.stderr, flush=True)

    if args.impl == 'reference':
        run_reference(args, rank, world, out)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    stamp('process group up')
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    pkg.lib.load()
    precision = args.precision
    if not pkg.lib.load().vqb_device_supports_tcgen05():
        raise SystemExit('bench.py: needs an sm_100 device')

    if args.ncu_step:
        run_workload(args.config, args, pkg, dev, rank, world, precision, args.steps, args.warmup, stamp, ncu_step=True)
        return
    head = run_workload(args.config, args, pkg, dev, rank, world, precision, args.steps, args.warmup, stamp, sample_clocks=True)

    subs = {}
    strict = None
    if args.config == 'cfg2' and not args.only_headline and not args.batch and not args.codebook and args.image_size == 256:
        def guarded(*a):                           # a failing sub-record must not take the headline line with it
            try:
                return run_workload(*a)
            except BaseException as e:                # incl. SystemExit from the loss check
                return {'error': f'{type(e).__name__}: {e}'[:400]}
        for name in ('cfg3', 'cfg4', 'cfg5'):
            subs[name] = guarded(name, args, pkg, dev, rank, world, precision, args.steps, args.warmup, stamp)
        if precision == 'fast':
            strict = guarded('cfg2', args, pkg, dev, rank, world, 'strict', min(args.steps, 3), 3, stamp)

    if rank == 0:
        pk, pk_kind = peaks()
        bs = head['config']['global_batch'] // world
        line = {
            'metric': head['metric'], 'value': head['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': head['dtype'], 'data': 'synthetic', 'config': head['config'], 'e2e': head['e2e'],
            'gpu_launches': head['gpu_launches'], 'gpu_launches_note': 'kernel entry points per timed region (the graph replays the same launches)',
            'cuda_graph': head.get('cuda_graph'), 'ms_per_step_eager': head.get('ms_per_step_eager'),
            'clocks': head.get('clocks'), 'roofline': head['roofline'], 'loss': head['loss'],
        }
        if world == 1:
            vq = {}
            for tag, (n_lat, k) in {'cfg2': (64 * 256, 1024), 'cfg5': (32 * 256, 8192)}.items():
                vq[tag] = vq_microbench(pkg, dev, n_lat, k, precision == 'fast', pk)
            line['vq_roofline'] = vq['cfg2']
            line['vq_roofline']['in_step_us_per_launch'] = head.get('vq_in_step_us_per_launch')
            line['vq_roofline_k8192'] = vq['cfg5']
            tr = os.path.join(ROOT, 'profiles', 'r02_vq_traffic.json')
            if os.path.exists(tr):
                try:
                    t = json.load(open(tr))
                    line['vq_roofline']['traffic'] = t.get('dram_bytes_per_launch')
                    line['vq_roofline']['traffic_source'] = t.get('source')
                    if isinstance(t.get('k8192'), dict):
                        line['vq_roofline_k8192']['traffic'] = t['k8192'].get('dram_bytes_per_launch')
                        line['vq_roofline_k8192']['traffic_source'] = t['k8192'].get('source')
                except Exception:
                    pass
        if subs:
            line['configs'] = subs
        if strict is not None:
            line['strict'] = strict
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            codebook = args.codebook or WORKLOADS[args.config]['codebook']
            try:
                ips, sec = cpu_reference_step(args.config, args.cpu_batch, args.image_size, codebook, args.cpu_steps, 1, threads)
                line['cpu_baseline'] = {'value': ips, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                        'sample': f'{args.cpu_steps} steps x {args.cpu_batch} images ({sec:.1f} s/step), oracle port of the '
                                                  f'reference modules, fp32 torch CPU'}
            except Exception as e:
                line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'error': str(e)[:300]}
        if world == 1 and not args.no_gpu_eager_baseline:
            codebook = args.codebook or WORKLOADS[args.config]['codebook']
            try:
                line['torch_eager_gpu_baseline'] = torch_eager_gpu_baseline(args.config, args.image_size, codebook, dev)
            except BaseException as e:
                line['torch_eager_gpu_baseline'] = {'error': f'{type(e).__name__}: {e}'[:300]}
            if isinstance(subs.get('cfg4'), dict) and 'error' not in subs['cfg4']:      # the VQGAN step (no R1 step in the timed region)
                try:
                    subs['cfg4']['torch_eager_gpu_baseline'] = torch_eager_gpu_baseline('cfg4', args.image_size, WORKLOADS['cfg4']['codebook'], dev)
                    subs['cfg4']['torch_eager_gpu_baseline']['note'] = (
                        'the discriminator\'s FIR resampling / bias-act run as the reference\'s pure-torch fallbacks (depthwise F.conv2d; '
                        'upfirdn2d.py:162-208, bias_act.py:55-97), not its CUDA plugin: pessimistic for the reference there; no R1 step timed')
                except BaseException as e:
                    subs['cfg4']['torch_eager_gpu_baseline'] = {'error': f'{type(e).__name__}: {e}'[:300]}
        out.write(json.dumps(line) + '\n'); out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
