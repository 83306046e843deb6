"""Time one VQGAN training step (BASELINE configs[3]-like: gumbel quantizer + LPIPS-VGG + StyleGAN2 discriminator active,
R1 on every step with --r1 1, off with --r1 0) through the public Trainer.run_step."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
from vqvae_vqgan_pytorch_lightning_b200.common_utils import derive_confs, get_model_conf
ap = argparse.ArgumentParser(); ap.add_argument('--batch', type=int, default=32); ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--r1', type=int, default=0); ap.add_argument('--conf', default='gumbel_vqgan'); ap.add_argument('--precision', default='fast')
a = ap.parse_args()
pkg.lib.load(); pkg.set_precision(a.precision)
conf = get_model_conf(os.path.join(ROOT, 'example_confs', a.conf + '.yaml'))
image_size, ae, q, l, t, bs = derive_confs(conf, 1, {'cumulative_bs': a.batch})
if l is not None and l.get('adversarial_params'):
    l = dict(l); l['adversarial_params'] = dict(l['adversarial_params'], start_epoch=0, r1_reg_weight=(10. if a.r1 else None), r1_reg_every=1)
torch.manual_seed(1234)
model = pkg.VQVAE(image_size, ae, q, l, t, pretrained_lpips=False).cuda().train()
tr = Trainer(); tr.attach(model); model.on_train_start(); model.training_augmentations = None
x = torch.rand(bs, 3, image_size, image_size, device='cuda')
for i in range(2): tr.run_step(x, i)
torch.cuda.synchronize()
names = list(pkg.lib.SIGNATURES.keys()); pkg.lib.timer = pkg.lib.KernelTimer(names)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps): loss = tr.run_step(x, 2 + i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
agg = {}
for name, args, s, e in pkg.lib.timer.records:
    key = name
    if name == 'vqb_conv2d_fwd': key = f'conv_fwd impl{args[0]} stride{args[16]}'
    if name == 'vqb_conv2d_wgrad': key = f'conv_wgrad impl{args[0]} stride{args[14]}'
    d = agg.setdefault(key, [0, 0.0]); d[0] += 1; d[1] += s.elapsed_time(e)
print(f'{a.conf} B={bs} {a.precision} r1={a.r1}: {ms:.1f} ms/step = {bs / ms * 1e3:.1f} img/s; loss {float(loss):.4f}; '
      f'peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB; launches/step {len(pkg.lib.timer.records) // a.steps}')
for k, (n, t_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f'{t_ / a.steps:9.2f} ms {n // a.steps:5d} calls  {k}')
# fp32 SIMT convolutions still on the path (impl 0): shape and time per call
seen = {}
for name, args, s, e in pkg.lib.timer.records:
    if name == 'vqb_conv2d_fwd' and args[0] == 0:
        key = ('fwd', args[9:17])            # N, H, W, Ci, Co, KH, KW, pad  (+ stride at 17)
    elif name == 'vqb_conv2d_wgrad' and args[0] == 0:
        key = ('wgrad', args[7:15])
    elif name == 'vqb_conv2d_dgrad':
        key = ('dgrad', args[5:14])
    else:
        continue
    d = seen.setdefault(key, [0, 0.0]); d[0] += 1; d[1] += s.elapsed_time(e)
for k, (n, t_) in sorted(seen.items(), key=lambda kv: -kv[1][1]):
    print(f'{t_ / a.steps:8.2f} ms {n // a.steps:3d} calls  {k}')
