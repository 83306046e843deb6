"""Forward 3x3 convolutions of the bench's main layer shapes (B = 64, bf16, tcgen05 path), with and without a fused residual:
device time per launch (CUDA events, median of 7) and algorithmic TFLOP/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
pkg.lib.load(); pkg.set_precision('fast')
B = int(os.environ.get('B', 64))
shapes = [(128, 128, 256), (128, 128, 128), (256, 256, 128), (256, 256, 64), (512, 512, 32), (512, 512, 16)]
cl = torch.channels_last
for ci, co, h in shapes:
    x = torch.randn(B, ci, h, h, device='cuda').bfloat16().contiguous(memory_format=cl)
    w = (torch.randn(co, ci, 3, 3, device='cuda') / (ci * 9) ** 0.5)
    res = torch.randn(B, co, h, h, device='cuda').bfloat16().contiguous(memory_format=cl)
    fl = 2.0 * B * h * h * ci * co * 9
    for name, r in (('plain', None), ('residual', res)):
        ts = []
        with torch.no_grad():
            for it in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); y = pkg.ops.conv2d(x, w, None, r, pad=1); e1.record(); torch.cuda.synchronize()
                if it >= 3: ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f'{ci:4d}->{co:4d} @{h:3d}^2 {name:9s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s')

# the narrow output head (decoder.conv_out 128 -> 3 + tanh)
x = torch.randn(B, 128, 256, 256, device='cuda').bfloat16().contiguous(memory_format=cl)
w = torch.randn(3, 128, 3, 3, device='cuda') * 0.03; b3 = torch.zeros(3, device='cuda')
ts = []
with torch.no_grad():
    for it in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = pkg.ops.conv2d(x, w, b3, None, pad=1, act=pkg.lib.ACT_TANH, out_dtype=torch.float32); e1.record(); torch.cuda.synchronize()
        if it >= 3: ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f' 128->   3 @256^2 head      {ms * 1e3:8.1f} us  {x.numel() * 2 / ms / 1e9:8.2f} TB/s of x')

# weight gradients of the same shapes
for ci, co, h in shapes:
    x = torch.randn(B, ci, h, h, device='cuda').bfloat16().contiguous(memory_format=cl).requires_grad_(False)
    w = (torch.randn(co, ci, 3, 3, device='cuda') / (ci * 9) ** 0.5).requires_grad_()
    dy = torch.randn(B, co, h, h, device='cuda').bfloat16().contiguous(memory_format=cl)
    fl = 2.0 * B * h * h * ci * co * 9
    pkg.lib.timer = pkg.lib.KernelTimer(['vqb_conv2d_wgrad'])
    for it in range(8):
        y = pkg.ops.conv2d(x, w, None, None, pad=1)
        y.backward(dy)
    torch.cuda.synchronize()
    ts = sorted(s_.elapsed_time(e_) for _, _, s_, e_ in pkg.lib.timer.records[3:])
    pkg.lib.timer = None
    ms = ts[len(ts) // 2]
    print(f'{ci:4d}->{co:4d} @{h:3d}^2 wgrad     {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s')
