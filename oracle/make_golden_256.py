"""Generate tests/golden/cfg2_256_ema.npz by EXECUTING THE REFERENCE'S OWN MODULES at the benchmark's image size and
architecture (test infrastructure only): ema_vqvae.yaml's autoencoder (channels 128, 2 ResBlocks per level, multipliers
1-2-2-4) and EMA quantizer (K = 1024, D = 256) on 256 x 256 images -- BASELINE configs[1] at batch 2, so that the fixture stays
small and the CPU oracle test finishes in seconds.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_256.py

Recipe: that of oracle/make_golden.py (torch.manual_seed(seed); quantizer -> Encoder -> Decoder as in vqvae/model.py:89-132;
init_codebook(); x = rand(B,3,S,S)*2-1; train mode; forward; (q_loss + mse).backward()), with ONE difference: the codebook (and
the EMA weight buffer) is then overwritten by a tie-free draw, torch.manual_seed(seed + 1); randn(K, D) * 0.05, because the
reference's U(+-1/K) initial codebook makes near-tied distances common and a tie would mask everything after the quantizer.
Full-size outputs are stored compactly: z whole (the latent is small), the reconstruction as its 8 x 8 average-pooled image plus
its first 4096 elements, gradients as per-tensor norms plus the two small head gradients."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, build          # noqa: E402  (the same reference-module construction)

CASE = dict(S=256, B=2, ch=128, nrb=2, mult=(1, 2, 2, 4), K=1024, D=256, seed=2468)
CODEBOOK_STD = 0.05


def main():
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    torch.set_num_threads(8)
    case = CASE
    torch.manual_seed(case['seed'])
    enc, q, dec = build(case, 'ema')
    x = torch.rand(case['B'], 3, case['S'], case['S']) * 2 - 1
    torch.manual_seed(case['seed'] + 1)
    with torch.no_grad():
        q.codebook.weight.copy_(torch.randn(case['K'], case['D']) * CODEBOOK_STD)
        q.ema_weight.copy_(q.codebook.weight)
    enc.train(); q.train(); dec.train()
    init_sum = {}
    for pre, m in (('encoder.', enc), ('quantizer.', q), ('decoder.', dec)):
        for n, t in m.state_dict().items():
            init_sum[pre + n] = float(t.double().abs().sum())
    z = enc(x)
    quant, idx, ql = q(z)
    recon = dec(quant)
    l2 = torch.nn.functional.mse_loss(recon, x)
    (ql + l2).backward()
    out = {
        'z': z.detach().numpy(), 'idx': idx.numpy(), 'q_loss': np.float32(ql.item()), 'l2': np.float32(l2.item()),
        'recon_pool8': torch.nn.functional.avg_pool2d(recon.detach(), 8).numpy(),
        'recon_head': recon.detach().reshape(-1)[:4096].numpy(),
        'recon_abs_sum': np.float64(recon.detach().double().abs().sum().item()),
        'grad_enc_conv_in': enc.conv_in.weight.grad.numpy(),
        'grad_dec_conv_out': dec.conv_out.weight.grad.numpy(),
        'grad_dec_conv_out_bias': dec.conv_out.bias.grad.numpy(),
        'new_ema_count': q.ema_count.numpy(),
        'new_ema_weight_rowsum': q.ema_weight.double().sum(1).numpy(),
        'new_codebook_rowsum': q.codebook.weight.detach().double().sum(1).numpy(),
    }
    names, norms = [], []
    for pre, m in (('encoder.', enc), ('quantizer.', q), ('decoder.', dec)):
        for n, t in m.named_parameters():
            if t.grad is not None:
                names.append(pre + n); norms.append(float(t.grad.double().norm()))
    out['grad_names'] = np.array(names)
    out['grad_norms'] = np.array(norms, dtype=np.float64)
    out['init_names'] = np.array(list(init_sum.keys()))
    out['init_abs_sums'] = np.array(list(init_sum.values()), dtype=np.float64)
    path = os.path.join(OUT, 'cfg2_256_ema.npz')
    np.savez_compressed(path, **out)
    print(f'{path}: q_loss={out["q_loss"]:.8f} l2={out["l2"]:.8f} sum(idx)={int(out["idx"].sum())} z.sum={out["z"].sum():.6f} '
          f'distinct codes={len(np.unique(out["idx"]))} ({os.path.getsize(path) / 1e6:.2f} MB)')


if __name__ == '__main__':
    main()
