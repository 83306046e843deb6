"""VQ kernel micro-benchmark: z ~ N(0,1) [N,256]; codebook U(+-1/K) (reference init, tie-heavy) and N(0,1)
(tie-free); exact fp32 SIMT kernel vs tensor-core search + exact re-evaluation.  Reports us/launch, algorithmic GB/s
(4ND + 4KD + 4ND + 8N + 8K + 12KD bytes, EMA statistics included) against the HBM peak, and the undecided-row count."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
pkg.lib.load()
PEAK = 6555.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
KERNELS = ['vqb_vq_fused', 'vqb_vq_assign', 'vqb_vq_assign_tc']
def timeit(fn, iters=20):
    """-> (median us of the whole op incl. output allocation / zero fills, median us of the search entry point alone)"""
    for _ in range(3): fn()
    ts, ks = [], []
    for _ in range(iters):
        flush.zero_()
        pkg.lib.timer = pkg.lib.KernelTimer(KERNELS)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
        ks.append(sum(v['ms'] for v in pkg.lib.timer.summary().values()) * 1e3)
        pkg.lib.timer = None
    ts.sort(); ks.sort(); return ts[len(ts) // 2], ks[len(ks) // 2]
rows = []
for (N, K) in ((16384, 1024), (8192, 8192)):
    D = 256
    for init in ('uniform', 'normal'):
        torch.manual_seed(0)
        z = torch.randn(N, D, device='cuda')
        cb = (torch.empty(K, D).uniform_(-1 / K, 1 / K) if init == 'uniform' else torch.randn(K, D)).cuda()
        nbytes = 4 * N * D * 2 + 4 * K * D + 8 * N + 8 * K + 12 * K * D
        prep = pkg.ops.CodebookPrep(); prep.get(cb)                        # cached split, as the quantizer modules hold it
        for tc in (('fused',) if '--fused-only' in sys.argv else (False, 'legacy', 'fused', 'fused+prep')):
            if tc == 'fused+prep':
                us, kus = timeit(lambda: pkg.ops.vq_assign_raw(z, cb, 0, True, True, use_tc='fused'))          # split launch included
            else:
                us, kus = timeit(lambda: pkg.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=tc, prep=prep if tc == 'fused' else None))
            und = int(pkg.ops.vq_assign_raw.last_undecided) if tc else N
            r = dict(N=N, K=K, init=init, kernel={False: 'fp32-simt', 'legacy': 'r01 tcgen05 search + exact rows (8 launches)',
                                                  'fused': 'vq_fused (1 launch, cached codebook split)',
                                                  'fused+prep': 'vq_fused + codebook split (2 launches)'}[tc], us=round(us, 1), kernel_us=round(kus, 1),
                     gbs=round(nbytes / us / 1e3, 1), frac_hbm=round(nbytes / us / 1e3 / PEAK, 4), undecided_rows=und,
                     tflops=round(2.0 * N * K * D * (3 if tc else 1) / us / 1e6, 1))
            rows.append(r); print(json.dumps(r), flush=True)
