import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from vqvae_vqgan_pytorch_lightning_b200.modules.loss import discriminator as D
from tests import common as C
pkg.lib.load(); pkg.set_precision('strict')
cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
g = C.golden('gan_discriminator')
torch.manual_seed(21)
d = D.Discriminator(64).cuda().train()
rec = {}
orig = D.Conv2dLayer.forward
def fwd(self, x, gain=1, residual=None):
    y = orig(self, x, gain, residual); y.retain_grad(); rec[id(self)] = y; return y
D.Conv2dLayer.forward = fwd
om = D.MinibatchStdLayer.forward
def mfwd(self, x):
    y = om(self, x); y.retain_grad(); rec['mb'] = y; return y
D.MinibatchStdLayer.forward = mfwd
img = cl(torch.from_numpy(g['img'])).requires_grad_()
logits = d(img); B = logits.shape[0]
w = torch.linspace(-1.0, 1.0, B, device='cuda').reshape(B, 1)
((logits * w).sum() + F.softplus(logits).mean()).backward()
# torch restatement (double precision ground truth on GPU)
f = torch.tensor([1., 3., 3., 1.], device='cuda', dtype=torch.float64); f = f.ger(f); f = f / f.sum()
def fir(x, pad, down):
    c = x.shape[1]
    return F.conv2d(F.pad(x, [pad] * 4), f[None, None].repeat(c, 1, 1, 1), groups=c)[:, :, ::down, ::down]
def layer(m, x, gain=1.0):
    wt = m.weight.detach().double() * m.weight_gain; b = m.bias.detach().double() if m.bias is not None else None
    if m.down == 1: y = F.conv2d(x, wt, None, padding=m.padding)
    elif m.kernel_size == 1: y = F.conv2d(fir(x, 1, 2), wt)
    else: y = F.conv2d(fir(x, m.padding + 1, 1), wt, stride=2)
    if b is not None: y = y + b.reshape(1, -1, 1, 1)
    if m.activation == 'lrelu': y = F.leaky_relu(y, 0.2) * (np.sqrt(2) * gain)
    else: y = y * gain
    y.retain_grad(); ref[id(m)] = y; return y
ref = {}
xi = torch.from_numpy(g['img']).cuda().double().requires_grad_()
x = None
for res in d.block_resolutions:
    b = getattr(d, f'b{res}')
    if b.in_channels == 0: x = layer(b.fromrgb, xi)
    yk = layer(b.skip, x, np.sqrt(0.5)); h = layer(b.conv0, x); h = layer(b.conv1, h, np.sqrt(0.5)); x = yk + h
N, Cc, H, W = x.shape
t = x.reshape(4, -1, 1, Cc, H, W); t = t - t.mean(dim=0); t = (t.square().mean(dim=0) + 1e-8).sqrt().mean(dim=[2, 3, 4])
x = torch.cat([x, t.reshape(-1, 1, 1, 1).repeat(4, 1, H, W)], dim=1)
x.retain_grad(); ref['mb'] = x
x = layer(d.b4.conv, x)
x = F.leaky_relu(x.flatten(1) @ (d.b4.fc.weight.detach().double() * d.b4.fc.weight_gain).t() + d.b4.fc.bias.detach().double(), 0.2) * np.sqrt(2)
lo = x @ (d.b4.out.weight.detach().double() * d.b4.out.weight_gain).t() + d.b4.out.bias.detach().double()
((lo * w.double()).sum() + F.softplus(lo).mean()).backward()
print('logits', C.rel_err(logits, lo), 'grad_img ours-vs-f64', C.rel_err(img.grad, xi.grad), 'fixture-vs-f64', C.rel_err(g['grad_img'], xi.grad))
for res in d.block_resolutions + [4]:
    b = getattr(d, f'b{res}')
    for name in ('fromrgb', 'skip', 'conv0', 'conv1', 'conv'):
        m = getattr(b, name, None)
        if m is None or id(m) not in rec: continue
        o, r = rec[id(m)], ref[id(m)]
        if name == 'skip':      # ours: skip output already includes conv1 (fused residual)
            r_out = ref[id(b.skip)] + ref[id(b.conv1)]; r_grad = ref[id(b.skip)].grad
        else: r_out, r_grad = r, r.grad
        print(f'b{res}.{name:8s} out rel {C.rel_err(o, r_out):.2e}  grad rel {C.rel_err(o.grad, r_grad):.2e}')

o, r = rec['mb'], ref['mb']
print('mbstd out rel', C.rel_err(o, r), 'grad rel (all ch)', C.rel_err(o.grad, r.grad), 'main ch', C.rel_err(o.grad[:, :512], r.grad[:, :512]), 'stat ch', C.rel_err(o.grad[:, 512:], r.grad[:, 512:]))
dd = (o.grad.double() - r.grad).abs(); print('max abs diff', float(dd.max()), 'at', torch.nonzero(dd == dd.max())[0].tolist(), 'ref val', float(r.grad.flatten()[dd.flatten().argmax()]), 'n>1e-3*max|g|', int((dd > 1e-3 * r.grad.abs().max()).sum()))
for res in d.block_resolutions + [4]:
    b = getattr(d, f'b{res}')
    for name in ('fromrgb', 'conv0', 'conv1', 'conv'):
        m = getattr(b, name, None)
        if m is None or id(m) not in rec: continue
        o, r = rec[id(m)], ref[id(m)]
        flips = int(((o > 0) != (r > 0)).sum())
        print(f'b{res}.{name}: sign flips {flips} of {o.numel()}  min|pre| ref {float(r.abs().min()):.2e}')
