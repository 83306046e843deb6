import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from vqvae_vqgan_pytorch_lightning_b200 import ops_gan
from tests import common as C
pkg.lib.load(); pkg.set_precision('strict')
cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
torch.manual_seed(0)
for (N, G, Cc, H, W, scale) in [(4, 4, 512, 4, 4, 1.0), (8, 4, 512, 4, 4, 1.0), (4, 4, 512, 4, 4, 0.01), (4, 4, 6, 4, 4, 1.0)]:
    x = (torch.randn(N, Cc, H, W) * scale).double(); xo = x.clone().requires_grad_()
    t = xo.reshape(G, -1, 1, Cc, H, W); t = t - t.mean(dim=0); t = (t.square().mean(dim=0) + 1e-8).sqrt().mean(dim=[2, 3, 4])
    y = torch.cat([xo, t.reshape(-1, 1, 1, 1).repeat(G, 1, H, W)], dim=1)
    go = torch.randn_like(y); go[:, :Cc] *= 1e-3          # emphasise the statistic's gradient
    y.backward(go)
    xg = cl(x.float()).requires_grad_()
    yg = ops_gan.mbstd(xg, G); yg.backward(cl(go.float()))
    print((N, G, Cc, H, W, scale), 'fwd', C.rel_err(yg, y), 'bwd', C.rel_err(xg.grad, xo.grad))
