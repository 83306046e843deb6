"""Per-shape table of the convolution launches inside one training step of a bench workload (eager, CUDA events around every
C-ABI entry point): calls, ms, algorithmic TFLOP/s per (direction, N, H, W, Ci, Co, k) -- which layers pull the conv roofline
down -- followed by the device time of every other entry point.

    python tools/conv_table.py --config cfg4 [--steps 2] [--json gpurun_out/conv_table_cfg4.json]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='cfg2'); ap.add_argument('--steps', type=int, default=2); ap.add_argument('--precision', default='fast')
ap.add_argument('--batch', type=int, default=0); ap.add_argument('--codebook', type=int, default=0); ap.add_argument('--image-size', type=int, default=256)
ap.add_argument('--json')
a = ap.parse_args()
pkg.lib.load(); pkg.set_precision(a.precision)
image_size, ae, q, l, t, bs, K = bench.model_confs(a.config, a, 1)
torch.manual_seed(1234)
model = pkg.VQVAE(image_size, ae, q, l, t, pretrained_lpips=False).cuda().train()
tr = Trainer(max_epochs=1, num_training_batches=64); tr.attach(model); model.on_train_start(); model.training_augmentations = None
x = torch.rand(bs, 3, image_size, image_size, device='cuda')
for i in range(3): tr.run_step(x, i)
torch.cuda.synchronize()
pkg.lib.timer = pkg.lib.KernelTimer(list(pkg.lib.SIGNATURES.keys()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps): tr.run_step(x, 3 + i)
e1.record(); torch.cuda.synchronize()
ms_step = e0.elapsed_time(e1) / a.steps
recs = pkg.lib.timer.records; pkg.lib.timer = None
conv, other = {}, {}
for name, args, s, e in recs:
    dt = s.elapsed_time(e)
    if name in bench.CONV_ENTRY_POINTS:
        sh = bench.conv_shape(name, args)
        d = conv.setdefault(sh[:-1], [0, 0.0, 0.0]); d[0] += 1; d[1] += dt; d[2] += sh[-1]
    else:
        d = other.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += dt
tot_ms = sum(v[1] for v in conv.values()) / a.steps
tot_fl = sum(v[2] for v in conv.values()) / a.steps
print(f'{a.config} B={bs} {a.precision}: {ms_step:.2f} ms/step eager; conv {tot_ms:.2f} ms, {tot_fl / 1e12:.2f} TFLOP/step -> {tot_fl / tot_ms / 1e9:.0f} TFLOP/s')
print(f'{"kind":14s} impl {"N":>4s} {"H":>4s} {"W":>4s} {"Ci":>5s} {"Co":>5s} k s res {"calls":>5s} {"ms/step":>8s} {"TFLOP/s":>8s} {"ms lost vs 1400":>8s}')
rows = []
for k, (n, ms, fl) in sorted(conv.items(), key=lambda kv: -kv[1][1]):
    tf = fl / ms / 1e9
    lost = (ms - fl / 1.4e12) / a.steps
    rows.append(dict(kind=k[0], impl=k[1], N=k[2], H=k[3], W=k[4], Ci=k[5], Co=k[6], k=k[7], stride=k[8], residual=k[9], calls=n // a.steps,
                     ms=ms / a.steps, tflops=tf, ms_lost=lost))
    print(f'{k[0]:14s} {k[1]:4d} {k[2]:4d} {k[3]:4d} {k[4]:4d} {k[5]:5d} {k[6]:5d} {k[7]} {k[8]} {int(k[9])}   {n // a.steps:5d} {ms / a.steps:8.3f} {tf:8.0f} {lost:8.3f}')
print('--- other entry points')
for k, (n, ms) in sorted(other.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f'{ms / a.steps:9.3f} ms {n // a.steps:5d} calls  {k}')
if a.json:
    json.dump({'config': a.config, 'batch': bs, 'ms_per_step_eager': ms_step, 'conv_ms': tot_ms, 'conv_tflop': tot_fl / 1e12, 'rows': rows,
               'other': {k: {'calls': n // a.steps, 'ms': ms / a.steps} for k, (n, ms) in other.items()}}, open(a.json, 'w'), indent=1)
