import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torchvision, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from vqvae_vqgan_pytorch_lightning_b200 import ops_gan, ops
from vqvae_vqgan_pytorch_lightning_b200.modules.loss.lpips import LPIPS
from tests import common as C
pkg.lib.load(); pkg.set_precision('strict')
cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
torch.manual_seed(0)
# (a) tap
fx, fy, w = torch.relu(torch.randn(2, 64, 8, 8)), torch.relu(torch.randn(2, 64, 8, 8)), torch.rand(1, 64, 1, 1)
fyo = fy.clone().requires_grad_()
nx = fx / (fx.pow(2).sum(1, keepdim=True).sqrt() + 1e-10); ny = fyo / (fyo.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
ref = F.conv2d((nx - ny) ** 2, w).mean((2, 3), True).sum(1).mean(); ref.backward()
fyg = cl(fy).requires_grad_(); out = ops_gan.lpips_tap(cl(fx), fyg, w.cuda()); out.backward()
print('tap fwd', float(out), float(ref), 'bwd rel', C.rel_err(fyg.grad, fyo.grad))
# (b) VGG trunk
tv = torchvision.models.vgg16(weights=None).features[:30].cuda()
m = LPIPS('vgg', pretrained=False); m.net.layers.load_state_dict(tv.state_dict()); m = m.cuda()
y = (torch.rand(2, 3, 64, 64) * 2 - 1)
yo = y.clone().cuda().requires_grad_()
mean, std = m.net.mean, m.net.std
h = (yo - mean) / std; taps = []
for i, l in enumerate(tv, 1):
    h = l(h) if not isinstance(l, torch.nn.ReLU) else F.relu(h)
    if i in (4, 9, 16, 23, 30): taps.append(h)
gs = [torch.randn_like(t) for t in taps]
sum((t * g).sum() for t, g in zip(taps, gs)).backward()
yg = cl(y).requires_grad_(); feats = m.net(yg)
sum((f * cl(g)).sum() for f, g in zip(feats, gs)).backward()
for i, (a, b) in enumerate(zip(feats, taps)): print('tap', i, 'fwd rel', C.rel_err(a, b))
print('trunk bwd rel', C.rel_err(yg.grad, yo.grad))
for k in range(5):
    yo.grad = None; yg.grad = None
    h = (yo - mean) / std; taps = []
    for i, l in enumerate(tv, 1):
        h = l(h) if not isinstance(l, torch.nn.ReLU) else F.relu(h)
        if i in (4, 9, 16, 23, 30): taps.append(h)
    (taps[k] * gs[k]).sum().backward()
    feats = m.net(yg); (feats[k] * cl(gs[k])).sum().backward()
    print('only tap', k, 'bwd rel', C.rel_err(yg.grad, yo.grad))
