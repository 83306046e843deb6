import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from tests import common as C
pkg.lib.load(); pkg.set_precision('strict')
cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
torch.manual_seed(0)
for (n, h, w, ci, co, k) in [(4, 4, 4, 513, 512, 3), (4, 4, 4, 512, 513, 3), (2, 8, 8, 130, 70, 3), (4, 4, 4, 512, 512, 3), (4,1,1,8192,512,1)]:
    x = torch.randn(n, ci, h, w).double(); wt = (torch.randn(co, ci, k, k) / (ci * k * k) ** 0.5).double(); b = torch.randn(co).double()
    xo, wo, bo = x.clone().requires_grad_(), wt.clone().requires_grad_(), b.clone().requires_grad_()
    y = F.leaky_relu(F.conv2d(xo, wo, bo, padding=k // 2), 0.2) * 1.4142135; go = torch.randn_like(y); y.backward(go)
    xg, wg, bg = cl(x.float()).requires_grad_(), wt.float().cuda().requires_grad_(), b.float().cuda().requires_grad_()
    yg = pkg.ops.conv2d(xg, wg, bg, None, pad=k // 2, act=pkg.lib.ACT_LRELU, alpha=0.2, gain=1.4142135); yg.backward(cl(go.float()))
    print((n, h, w, ci, co, k), 'y', C.rel_err(yg, y), 'dx', C.rel_err(xg.grad, xo.grad), 'dw', C.rel_err(wg.grad, wo.grad), 'db', C.rel_err(bg.grad, bo.grad))
