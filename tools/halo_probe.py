"""Probe which UMMA descriptor convention makes the 3x3 halo-reuse kernel exact (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from tests import common as C
L = pkg.lib.load(); pkg.set_precision('fast')
torch.manual_seed(0)
for (n, h, w, ci, co) in [(2, 32, 32, 128, 128), (1, 16, 24, 64, 256), (3, 64, 64, 128, 128)]:
    x = torch.randn(n, ci, h, w).bfloat16().float(); wt = (torch.randn(co, ci, 3, 3) / (ci * 9) ** 0.5).bfloat16().float()
    ref = F.conv2d(x, wt, None, padding=1)
    xg = x.cuda().contiguous(memory_format=torch.channels_last).bfloat16(); wg = wt.cuda()
    for mode in (0, 1, 2, 3, 4):
        L.vqb_set_halo_mode(mode)
        y = pkg.ops.conv2d(xg, wg, None, None, pad=1, out_dtype=torch.float32)
        torch.cuda.synchronize()
        print(f'shape {(n, h, w, ci, co)} mode {mode}: rel_err {C.rel_err(y, ref):.3e}', flush=True)
L.vqb_set_halo_mode(-1)
