"""print the error components of the fast stride-2 conv route against fp32 torch"""
import numpy as np, torch, torch.nn.functional as F
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from vqvae_vqgan_pytorch_lightning_b200.modules.loss.discriminator import Conv2dLayer, setup_filter
from tests import common as C
pkg.lib.load(); pkg.set_precision('fast')
r16 = lambda t: t.bfloat16().float()
cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
torch.manual_seed(4)
n, ci, co, h = 2, 128, 256, 32
layer = Conv2dLayer(ci, co, kernel_size=3, activation='lrelu', down=2)
with torch.no_grad():
    layer.weight.copy_(r16(layer.weight)); layer.bias.copy_(torch.randn(co) * 0.1)
x = r16(torch.randn(n, ci, h, h)); go = r16(torch.randn(n, co, h // 2, h // 2))
f = setup_filter()
for rnd in (False, True):
    xo = x.clone().requires_grad_(); wo = layer.weight.detach().clone().cpu().requires_grad_(); bo = layer.bias.detach().clone().cpu().requires_grad_()
    xb = F.conv2d(F.pad(xo, [2, 2, 2, 2]), f[None, None].repeat(ci, 1, 1, 1), groups=ci)
    if rnd:
        xb = xb + (r16(xb) - xb).detach()
    wq = wo * layer.weight_gain
    if rnd:
        wq = wq + (r16(wq) - wq).detach()
    pre = F.conv2d(xb, wq, bo, stride=2)
    pre.retain_grad()
    yo = F.leaky_relu(pre, 0.2) * (np.sqrt(2) * np.sqrt(0.5))
    yo.backward(go)
    L = layer.cuda(); L.zero_grad()
    xg = cl(x).bfloat16().requires_grad_()
    yg = L(xg, gain=np.sqrt(0.5))
    yg.backward(cl(go).bfloat16())
    flips = ((yg.float().cpu() > 0) != (yo > 0)).float().mean().item()
    print('round_ref', rnd, 'y', C.rel_err(yg.float(), yo), 'dx', C.rel_err(xg.grad.float(), xo.grad), 'dw', C.rel_err(L.weight.grad, wo.grad),
          'db', C.rel_err(L.bias.grad, bo.grad), 'sign flips', flips)
