"""How the fused VQ search behaves on the codebook a TRAINING run produces (not on a synthetic one): runs the cfg5 step
(EMA VQGAN, K = 8192, B = 32) eagerly for --steps steps and prints, every 10th step, the code-norm distribution, the number of
rows that took the exact re-rank / the exact full scan, and the duration of the VQ launch.
(round 2: after ~80 steps the unused codes of the EMA codebook have decayed towards the origin and every row overflowed its
candidate list under the global-norm error band -- 30 ms per launch; per-code bounds keep it at the isolated figure.)"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=140)
    ap.add_argument('--config', default='cfg5')
    a = ap.parse_args()
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    args = argparse.Namespace(batch=0, codebook=0, image_size=256)
    pkg.set_precision('fast')
    dev = torch.device('cuda:0')
    image_size, ae_conf, q_conf, l_conf, t_conf, bs, codebook = bench.model_confs(a.config, args, 1)
    torch.manual_seed(1234)
    model = pkg.VQVAE(image_size, ae_conf, q_conf, l_conf, t_conf, pretrained_lpips=False).to(dev).train()
    trainer = Trainer(max_epochs=1, num_training_batches=a.steps + 4, cuda_graph=False)
    trainer.attach(model)
    model.on_train_start()
    model.training_augmentations = None
    xs = [torch.rand(bs, 3, image_size, image_size, device=dev) for _ in range(2)]
    for i in range(a.steps):
        rec = i % 10 == 9
        if rec:
            pkg.lib.timer = pkg.lib.KernelTimer(['vqb_vq_fused'])
        trainer.run_step(xs[i % 2], i)
        if rec:
            torch.cuda.synchronize()
            ks = pkg.lib.timer.summary().get('vqb_vq_fused')
            pkg.lib.timer = None
            cb = model.quantizer.codebook.weight.detach()
            nrm = cb.norm(dim=1)
            qs = torch.quantile(nrm, torch.tensor([0.0, 0.5, 0.9, 0.99, 1.0], device=dev)).tolist()
            und = int(pkg.ops.vq_assign_raw.last_undecided)
            full = int(pkg.ops.vq_assign_raw.last_fullscan)
            us = ks['ms'] / ks['calls'] * 1e3 if ks else float('nan')
            print(f'step {i + 1:4d}: vq {us:9.1f} us  re-ranked {und:5d}  full-scan {full:5d}  |e| min/med/p90/p99/max '
                  + ' '.join(f'{v:.3g}' for v in qs), flush=True)


if __name__ == '__main__':
    main()
