"""GroupNorm+SiLU forward/backward on the largest layer shape of the bench (64 x 128 x 256 x 256, bf16): device time per
kernel via CUDA events (KernelTimer), or run under `ncu --set full -k regex:gn_` for stall analysis."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
pkg.lib.load(); pkg.set_precision('fast')
n, c, h = int(os.environ.get('GN_N', 64)), int(os.environ.get('GN_C', 128)), int(os.environ.get('GN_H', 256))
x = torch.randn(n, c, h, h, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_()
g = torch.ones(1, c, 1, 1, device='cuda', requires_grad=True); b = torch.zeros(1, c, 1, 1, device='cuda', requires_grad=True)
dy = torch.randn_like(x); dskip = torch.randn_like(x)
T = x.numel() * 2
for it in range(3):
    if it == 2:
        pkg.lib.timer = pkg.lib.KernelTimer([k for k in pkg.lib.SIGNATURES if k.startswith('vqb_gn')])
    y, skip = pkg.ops.group_norm_act(x, g, b, 32, 1e-6, pkg.lib.ACT_SILU, True)
    torch.autograd.backward([y, skip], [dy, dskip])
torch.cuda.synchronize()
traffic = {'vqb_gn_stats': 1, 'vqb_gn_apply': 2, 'vqb_gn_bwd_reduce': 2, 'vqb_gn_bwd_apply': 4}
for name, args, s, e in pkg.lib.timer.records:
    ms = s.elapsed_time(e)
    k = traffic.get(name)
    print(f'{name:22s} {ms * 1e3:9.1f} us' + (f'  {k * T / ms / 1e9:7.2f} TB/s ({k}T)' if k else ''))
