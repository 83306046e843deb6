"""Where does the end-to-end (host buffers, per-step loss read) step lose time against the device-resident loop?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
from vqvae_vqgan_pytorch_lightning_b200.common_utils import derive_confs, get_model_conf
pkg.lib.load(); pkg.set_precision('fast')
conf = get_model_conf(os.path.join(ROOT, 'example_confs', 'ema_vqvae.yaml'))
image_size, ae, q, l, t, bs = derive_confs(conf, 1, {'num_embeddings': 1024, 'cumulative_bs': 64})
for graph in (False, True):
    torch.manual_seed(1234)
    model = pkg.VQVAE(image_size, ae, q, l, t).cuda().train()
    tr = Trainer(max_epochs=1, num_training_batches=200, cuda_graph=graph); tr.attach(model); model.on_train_start(); model.training_augmentations = None
    host = [torch.rand(bs, 3, image_size, image_size).pin_memory() for _ in range(2)]
    res = [h.cuda() for h in host]
    for i in range(6): tr.run_step(res[i % 2], i)
    torch.cuda.synchronize()
    def loop(kind, n=8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(n):
            if kind == 'resident': loss = tr.run_step(res[i % 2], 10 + i)
            elif kind == 'resident+sync': loss = tr.run_step(res[i % 2], 10 + i); float(loss)
            elif kind == 'h2d': loss = tr.run_step(host[i % 2].to('cuda', non_blocking=True), 10 + i)
            elif kind == 'h2d+sync': loss = tr.run_step(host[i % 2].to('cuda', non_blocking=True), 10 + i); float(loss.detach().cpu())
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
    for kind in ('resident', 'resident+sync', 'h2d', 'h2d+sync', 'resident'):
        print(f'graph={graph} {kind:14s} {loop(kind):7.2f} ms/step', flush=True)
    del model, tr; torch.cuda.empty_cache()
