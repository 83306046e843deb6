"""Per-entry-point / per-shape device-time breakdown of one training step (CUDA events around every C-ABI call)."""
import argparse, collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
from vqvae_vqgan_pytorch_lightning_b200.common_utils import derive_confs, get_model_conf

ap = argparse.ArgumentParser()
ap.add_argument('--precision', default='fast'); ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--steps', type=int, default=2); ap.add_argument('--top', type=int, default=40)
a = ap.parse_args()
pkg.lib.load(); pkg.set_precision(a.precision)
conf = get_model_conf(os.path.join(ROOT, 'example_confs', 'ema_vqvae.yaml'))
image_size, ae, q, l, t, bs = derive_confs(conf, 1, {'num_embeddings': 1024, 'cumulative_bs': a.batch})
torch.manual_seed(1234)
model = pkg.VQVAE(image_size, ae, q, l, t).cuda().train()
tr = Trainer(); tr.attach(model); model.on_train_start(); model.training_augmentations = None
x = torch.rand(bs, 3, image_size, image_size, device='cuda')
for i in range(2):
    tr.run_step(x, i)
torch.cuda.synchronize()
names = list(pkg.lib.SIGNATURES.keys())
pkg.lib.timer = pkg.lib.KernelTimer(names)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    tr.run_step(x, 2 + i)
e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1) / a.steps
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for name, args, s, e in pkg.lib.timer.records:
    ms = s.elapsed_time(e)
    key = name; fl = 0.0
    if name in ('vqb_conv2d_fwd', 'vqb_conv2d_fwd_gn'):
        impl = args[0]; n, h, w, ci, co, kh, kw, pad, st = args[8:17]
        key = f'conv_fwd{"+gn" if name.endswith("gn") else ""} impl{impl} {ci}->{co} k{kh} @{h}x{w}'; fl = 2.0 * n * h * w * ci * co * kh * kw
    elif name == 'vqb_conv2d_wgrad':
        impl = args[0]; n, h, w, ci, co, kh, kw, pad, st = args[6:15]
        key = f'conv_wgrad impl{impl} {ci}->{co} k{kh} @{h}x{w}'; fl = 2.0 * n * h * w * ci * co * kh * kw
    elif name.startswith('vqb_gn_'):
        key = f'{name} C={args[-4] if name!="vqb_gn_finalize" else ""}'
    r = agg[key]; r[0] += 1; r[1] += ms; r[2] += fl
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
print(f'step {total:.2f} ms; sum of timed calls {sum(r[1] for _, r in rows) / a.steps:.2f} ms/step')
for k, (cnt, ms, fl) in rows[:a.top]:
    tf = fl / (ms / 1e3) / 1e12 if fl else 0
    print(f'{ms / a.steps:9.3f} ms  {cnt // a.steps:4d} calls  {tf:7.1f} TF/s  {k}')
# ---- CPU enqueue time per step (no device sync inside) ----
import time
pkg.lib.timer = None
torch.cuda.synchronize()
ts = []
for i in range(3):
    t0 = time.perf_counter(); tr.run_step(x, 10 + i); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
print('CPU enqueue ms / total ms per step:', ['%.1f / %.1f' % t for t in ts], 'launches/step', pkg.lib.launch_count // (2 + a.steps + 3))
