#!/usr/bin/env python
"""bench.py -- images/sec of one full VQ-VAE train step (forward + backward + AdamW) at 256x256.

    python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N ...          # the reference's CPU arithmetic (oracle port) on host cores

Workload (BASELINE.json configs[1]): example_confs/ema_vqvae.yaml with codebook 1024, 256x256 synthetic RGB,
batch 64 per GPU (weak scaling), random-init weights in the reference's construction order, augmentation off.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
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

METRIC = 'images/sec VQ-VAE train step 256x256 (ema_vqvae, K=1024)'
UNIT = 'images/s'
FALLBACK_PEAKS = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}


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


def model_confs(args, world):
    from vqvae_vqgan_pytorch_lightning_b200.common_utils import derive_confs, get_model_conf
    conf = get_model_conf(os.path.join(ROOT, 'example_confs', 'ema_vqvae.yaml'))
    return derive_confs(conf, world, {'num_embeddings': args.codebook, 'image_size': args.image_size,
                                      'cumulative_bs': args.batch * world})


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

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def conv_flops(name, a):
    if name == 'vqb_conv2d_fwd':
        n, h, w, ci, co, kh, kw, pad, stride = a[8:17]
    else:
        n, h, w, ci, co, kh, kw, pad, stride = a[6:15]
    oh, ow = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    return 2.0 * n * oh * ow * co * ci * kh * kw


# ------------------------------------------------------------------------------------------------------
def vq_microbench(pkg, dev, n_lat, k, use_tc, hbm_peak, iters=15):
    """The fused VQ kernel in isolation on the workload's latent shape: z ~ N(0,1) [N,256]; codebook N(0,1)
    (tie-free, what a trained codebook looks like to the search) and U(+-1/K) (the reference's initial codebook: thousands of
    codes within 1e-5 of each other, so the exact fp32 path takes over).  L2 is flushed between launches; algorithmic bytes =
    4ND (z) + 4KD (codebook) + 4ND (q) + 8N (idx) + 8K + 12KD (EMA statistics)."""
    d = 256
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    nbytes = 4 * n_lat * d * 2 + 4 * k * d + 8 * n_lat + 8 * k + 12 * k * d
    cases = []
    for init in ('normal', 'uniform'):
        torch.manual_seed(0)
        z = torch.randn(n_lat, d, device=dev)
        cb = torch.randn(k, d, device=dev) if init == 'normal' else torch.empty(k, d, device=dev).uniform_(-1 / k, 1 / k)
        for _ in range(3):
            pkg.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=use_tc)
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); pkg.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=use_tc); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        und = int(pkg.ops.vq_assign_raw.last_undecided) if use_tc else None
        cases.append({'codebook': init, 'us_per_launch': us, 'gbs': nbytes / us / 1e3, 'frac': nbytes / us / 1e3 / hbm_peak,
                      'rows_on_exact_path': und})
    c0 = cases[0]
    return {'bound': 'hbm', 'kernel': 'vq_assign_tc: bf16 hi/lo tcgen05 search + exact fp32 re-rank + gather/EMA scatter' if use_tc
            else 'vq_assign: fused fp32 SIMT', 'achieved': c0['gbs'], 'peak': hbm_peak, 'unit': 'GB/s', 'frac': c0['frac'],
            'traffic': None, 'algorithmic_bytes': nbytes, 'shape': {'N': n_lat, 'K': k, 'D': d}, 'cases': cases,
            'note': 'exact-argmin contract: the distance GEMM (2NKD x3 bf16-split FLOP) keeps this kernel tensor/latency-bound, not HBM-bound'}


def cpu_reference_step(batch: int, image_size: int, codebook: int, steps: int, warmup: int, threads: int):
    """The reference's arithmetic on host cores: oracle port of forward + backward + AdamW (fp32, oneDNN/MKL).
    Returns (images_per_sec, seconds_per_step)."""
    from oracle import init_state as oinit
    from oracle import vqvae_oracle as orc
    torch.set_num_threads(threads)
    sd = oinit.init_state('ema', codebook, 256, 128, 2, (1, 2, 2, 4), seed=1234)
    sd = oinit.make_leaf(sd, 'ema')
    cfg = {'num_res_blocks': 2, 'channel_multipliers': (1, 2, 2, 4),
           'quantizer': dict(type='ema', commitment_cost=0.25, decay=0.95, epsilon=1e-5)}
    mom = {n: (torch.zeros_like(t), torch.zeros_like(t)) for n, t in sd.items() if t.requires_grad}
    x = orc.normalize_images(torch.rand(batch, 3, image_size, image_size))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for t in sd.values():
            t.grad = None
        out = orc.train_step_mse(sd, x, cfg)
        with torch.no_grad():
            sd['quantizer.codebook.weight'].copy_(out['new_codebook'])
            sd['quantizer.ema_count'].copy_(out['new_ema_count'])
            sd['quantizer.ema_weight'].copy_(out['new_ema_weight'])
            for n, t in sd.items():
                if t.requires_grad and t.grad is not None:
                    orc.adamw_step(t, t.grad, mom[n][0], mom[n][1], it + 1, 1e-4, 0.0, 0.99, 1e-8, 1e-4)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch / sec, sec


def run_reference(args, rank, world, out=sys.stdout):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    b = args.cpu_batch
    ips, sec = cpu_reference_step(b, args.image_size, args.codebook, args.steps, min(args.warmup, 1), threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': min(args.warmup, 1), 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'ema_vqvae {args.image_size}x{args.image_size} K={args.codebook}: bounded sample of the '
                               f'B={args.batch} step at micro-batch {b} (img/s is batch-insensitive on CPU)'},
        'cpu_baseline': {'value': ips, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': f'{args.steps} steps x {b} images, oracle port of the reference modules (fp32, torch CPU)'},
        'e2e': {'value': ips, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    out.write(json.dumps(line) + '\n'); out.flush()


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('VQB_PRECISION', 'fast'), choices=['strict', 'fast'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step')
    ap.add_argument('--image-size', type=int, default=256)
    ap.add_argument('--codebook', type=int, default=1024)
    ap.add_argument('--cpu-batch', type=int, default=2)
    ap.add_argument('--cpu-steps', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    pkg.lib.load()
    precision = args.precision
    if precision == 'fast' and not pkg.lib.load().vqb_device_supports_tcgen05():
        raise SystemExit('bench.py: fast precision needs an sm_100 device')
    pkg.set_precision(precision)

    image_size, ae_conf, q_conf, l_conf, t_conf, bs = model_confs(args, world)
    torch.manual_seed(1234)                       # same weights on every rank (reference construction order)
    model = pkg.VQVAE(image_size, ae_conf, q_conf, l_conf, t_conf).to(dev).train()
    trainer = Trainer(max_epochs=1, num_training_batches=args.steps + args.warmup)
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

    def step_resident(i):
        loss = trainer.run_step(resident[i % nbuf], step_idx[0]); step_idx[0] += 1
        return loss

    def step_e2e(i):
        x = host[i % nbuf].to(dev, non_blocking=True)                            # H2D from pinned memory, every step
        loss = trainer.run_step(x, step_idx[0]); step_idx[0] += 1
        return float(loss.detach().cpu())                                        # D2H read of the step's loss

    stamp('model and inputs ready')
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    stamp('warm-up done')
    if args.ncu_step:
        torch.cuda.profiler.start()
        step_resident(0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # ---- timed region 1: inputs resident in HBM (value, ms_per_step, roofline) ---------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    pkg.lib.timer = pkg.lib.KernelTimer(['vqb_conv2d_fwd', 'vqb_conv2d_wgrad', 'vqb_vq_assign', 'vqb_vq_assign_tc'])
    launches0 = pkg.lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pkg.lib.launch_count - launches0
    ksum = pkg.lib.timer.summary()
    pkg.lib.timer = None

    # ---- timed region 2: end to end through the public API with host buffers ---------------------------------------
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record()
    for i in range(args.steps):
        step_e2e(i)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    stamp('timed regions done')
    clk = clocks.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        pk, pk_kind = peaks()
        total_images = bs * world * args.steps
        value = total_images / (ms / 1e3)
        e2e = total_images / (ms_e2e / 1e3)
        conv_ms = sum(ksum[k]['ms'] for k in ('vqb_conv2d_fwd', 'vqb_conv2d_wgrad') if k in ksum)
        conv_fl = sum(conv_flops(k, a) for k in ('vqb_conv2d_fwd', 'vqb_conv2d_wgrad') if k in ksum for a in ksum[k]['args'])
        conv_calls = sum(ksum[k]['calls'] for k in ('vqb_conv2d_fwd', 'vqb_conv2d_wgrad') if k in ksum)
        achieved = conv_fl / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        peak = pk['bf16_tflops_sustained']
        roofline = {'bound': 'tensor', 'kernel': 'implicit-GEMM conv (fwd+dgrad+wgrad launches)', 'achieved': achieved, 'peak': peak,
                    'unit': 'TFLOP/s', 'frac': achieved / peak,
                    # DRAM bytes of ONE launch of the largest layer (128->128 3x3 @256^2, B=64; algorithmic 2.147e9 B = x + y) from the
                    # committed `ncu --set full` capture -- not measured live (a profiler cannot run inside the timed region)
                    'traffic': 2.114e9 if (image_size, bs) == (256, 64) else None,
                    'traffic_source': 'profiles/r01_ncu_full_top_kernels.txt: conv_fwd_tc_halo_t_kernel launch 0, dram read 1.084 GB + write 1.030 GB',
                    'peak_source': f'{pk_kind} (sustained bf16)',
                    'launches_timed': conv_calls, 'share_of_step': conv_ms / ms, 'flop_per_step': conv_fl / args.steps}
        vq_roof = vq_microbench(pkg, dev, bs * (image_size // 16) ** 2, args.codebook, precision == 'fast', pk['hbm_gbs'])
        vq_step = ksum.get('vqb_vq_assign_tc') or ksum.get('vqb_vq_assign')
        if vq_step:
            vq_roof['in_step_us_per_launch'] = vq_step['ms'] / vq_step['calls'] * 1e3      # init-time codebook: tie-heavy
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if precision == 'fast' else 'f32', 'data': 'synthetic',
            'config': {'workload': f'ema_vqvae {image_size}x{image_size} K={args.codebook} B={bs}/GPU (BASELINE configs[1]), '
                                   f'fwd+bwd+AdamW, augmentation off', 'precision': precision, 'global_batch': bs * world,
                       'parallelism': f'dp{world}', 'l2_policy': 'working set per step (inputs 50 MB + activations > 10 GB) exceeds the 126 MB L2'},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': bs * 3 * image_size * image_size * 4, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'clocks': clk, 'roofline': roofline, 'vq_roofline': vq_roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            ips, sec = cpu_reference_step(args.cpu_batch, image_size, args.codebook, args.cpu_steps, 1, threads)
            line['cpu_baseline'] = {'value': ips, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                    'sample': f'{args.cpu_steps} steps x {args.cpu_batch} images ({sec:.1f} s/step), oracle port of the '
                                              f'reference modules, fp32 torch CPU'}
        out.write(json.dumps(line) + '\n'); out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
