import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from vqvae_vqgan_pytorch_lightning_b200.modules.loss.discriminator import Discriminator
from tests import common as C
pkg.lib.load(); pkg.set_precision('strict')
cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
g = C.golden('gan_discriminator')
torch.manual_seed(21)
d = Discriminator(64).cuda().train()
img = cl(torch.from_numpy(g['img'])).requires_grad_()
logits = d(img); B = logits.shape[0]
w = torch.linspace(-1.0, 1.0, B, device='cuda').reshape(B, 1)
((logits * w).sum() + F.softplus(logits).mean()).backward()
print('grad_img', C.rel_err(img.grad, g['grad_img']))
ref = dict(zip(g['grad_names'].tolist(), g['grad_norms'].tolist()))
for n, p in d.named_parameters():
    print(f'{n:28s} ours {float(p.grad.double().norm()):.6e} ref {ref[n]:.6e} rel {abs(float(p.grad.double().norm())-ref[n])/ref[n]:.2e}')
