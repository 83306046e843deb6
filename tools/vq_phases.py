"""Phase timeline of the fused VQ kernel (vqb_vq_fused_set_trace: globaltimer stamps per CTA at the phase boundaries).
   python tools/vq_phases.py [N K init]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
L = pkg.lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
init = sys.argv[3] if len(sys.argv) > 3 else 'normal'
torch.manual_seed(0)
z = torch.randn(N, 256, device='cuda')
cb = (torch.empty(K, 256).uniform_(-1 / K, 1 / K) if init == 'uniform' else torch.randn(K, 256)).cuda()
prep = pkg.ops.CodebookPrep(); prep.get(cb)
ctas = 2 * ((N + 255) // 256)
trace = torch.zeros(ctas, 8, dtype=torch.int64, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3):
    pkg.ops.vq_assign_raw(z, cb, 0, True, True, use_tc='fused', prep=prep)
names = ['prologue (z -> fp16 smem, norms)', 'scan (MMA + candidate lists)', 'decide', 'exact re-rank', 'finish (gather, q, sse, EMA sums)']
for cold in (False, True):
    if cold: flush.zero_()
    L.vqb_vq_fused_set_trace(ctypes.c_void_p(trace.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pkg.ops.vq_assign_raw(z, cb, 0, True, True, use_tc='fused', prep=prep); e1.record()
    torch.cuda.synchronize()
    L.vqb_vq_fused_set_trace(None)
    t = trace.cpu().double()
    t0 = t[:, 0].min()
    print(f'N={N} K={K} {init} {"cold L2" if cold else "warm L2"}: {e0.elapsed_time(e1) * 1e3:.1f} us by events; CTAs {ctas}; first CTA start .. last CTA end '
          f'{(t[:, 5].max() - t0) / 1e3:.1f} us; CTA start skew {(t[:, 0].max() - t0) / 1e3:.1f} us; undecided rows {int(pkg.ops.vq_assign_raw.last_undecided)}')
    for i, nm in enumerate(names):
        d = (t[:, i + 1] - t[:, i]) / 1e3
        print(f'   {nm:42s} mean {d.mean():6.2f} us   max {d.max():6.2f} us   (ends at mean +{(t[:, i + 1] - t0).mean() / 1e3:6.2f} us)')
