"""Driver for `ncu -k regex:vq_fused`: a few launches of the fused VQ kernel at one bench shape.
   python tools/vq_profile.py [N K init]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
pkg.lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
init = sys.argv[3] if len(sys.argv) > 3 else 'normal'
torch.manual_seed(0)
z = torch.randn(N, 256, device='cuda')
cb = (torch.empty(K, 256).uniform_(-1 / K, 1 / K) if init == 'uniform' else torch.randn(K, 256)).cuda()
prep = pkg.ops.CodebookPrep(); prep.get(cb)
for _ in range(4):
    pkg.ops.vq_assign_raw(z, cb, 0, True, True, use_tc='fused', prep=prep)
torch.cuda.synchronize()
print('undecided', int(pkg.ops.vq_assign_raw.last_undecided))
