"""Run the 2-CTA UMMA probe (tools/probes/umma2_probe.cu) against torch.  ALWAYS run under `timeout`: a protocol error
leaves the epilogue warps spinning on an mbarrier."""
import ctypes, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
so = os.path.join(HERE, 'libumma2_probe.so')
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, 'umma2_probe.cu')):
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O2', '-std=c++17', '-shared', '-Xcompiler', '-fPIC',
                           '-I' + os.path.join(ROOT, 'vqvae_vqgan_pytorch_lightning_b200', 'csrc'), os.path.join(HERE, 'umma2_probe.cu'),
                           '-o', so, '-lcudart'])
import torch
lib = ctypes.CDLL(so)
lib.umma2_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
torch.manual_seed(0)
for K in (64, 256):
    A = torch.randn(256, K, device='cuda').bfloat16(); B = torch.randn(256, K, device='cuda').bfloat16()
    D = torch.full((256, 256), float('nan'), device='cuda'); dbg = torch.full((2,), -1, dtype=torch.int32, device='cuda')
    rc = lib.umma2_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, dbg.data_ptr(), torch.cuda.current_stream().cuda_stream)
    print('K', K, 'launch rc', rc, flush=True)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (D - ref).abs().max().item()
    print('K', K, 'tmem bases', dbg.tolist(), 'max abs err', err, 'ref max', ref.abs().max().item(),
          'rows 0-127 err', (D[:128] - ref[:128]).abs().max().item(), 'rows 128-255 err', (D[128:] - ref[128:]).abs().max().item(),
          'cols 0-127 err', (D[:, :128] - ref[:, :128]).abs().max().item(), 'cols 128-255 err', (D[:, 128:] - ref[:, 128:]).abs().max().item(), flush=True)
