"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN MODULES (test infrastructure only).

Run in the build container (the reference is mounted read-only at /root/reference and cannot travel
to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Recipe: torch.manual_seed(seed); build quantizer -> Encoder -> Decoder in the order of
vqvae/model.py:89-132; quantizer.init_codebook(); x = rand(B,3,S,S)*2-1; train mode; forward;
(q_loss + mse(recon, x)).backward().  Everything the fixtures hold is an OUTPUT of reference code; the
inputs are regenerated from the seed by the tests (initial-weight checksums are stored to detect drift).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = os.environ.get('VQ_REF_PATH', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

CASES = {
    # name: (image, batch, channels, num_res_blocks, multipliers, K, D)
    'cfg1': dict(S=64, B=8, ch=128, nrb=2, mult=(1, 2, 2, 4), K=256, D=256, seed=1234),
    'tiny': dict(S=16, B=4, ch=32, nrb=1, mult=(1, 2), K=64, D=32, seed=4321),
}
Q_PARAMS = {
    'standard': dict(commitment_cost=0.25),
    'ema': dict(commitment_cost=0.25, decay=0.95, epsilon=1e-5),
    'entropy': dict(ent_loss_ratio=0.1, ent_temperature=0.01, ent_loss_type='softmax', commitment_cost=0.25),
    'gumbel': dict(straight_through=False, temp=1.0, kl_cost=0.00859375),
}


def build(case: dict, qtype: str):
    from vqvae.modules.autoencoder import Encoder, Decoder
    from vqvae.modules import vector_quantizers as vq
    K, D = case['K'], case['D']
    p = Q_PARAMS[qtype]
    if qtype == 'standard':
        q = vq.VectorQuantizer(K, D, p['commitment_cost'])
    elif qtype == 'ema':
        q = vq.EMAVectorQuantizer(K, D, p['commitment_cost'], p['decay'], p['epsilon'])
    elif qtype == 'entropy':
        q = vq.EntropyVectorQuantizer(K, D, p['ent_loss_ratio'], p['ent_temperature'], p['ent_loss_type'],
                                      p['commitment_cost'])
    else:
        q = vq.GumbelVectorQuantizer(K, D, p['straight_through'], p['temp'], p['kl_cost'])
    enc = Encoder(case['ch'], case['nrb'], case['mult'], K if qtype == 'gumbel' else D)
    dec = Decoder(case['ch'], case['nrb'], case['mult'], D)
    q.init_codebook()
    return enc, q, dec


def run_case(name: str, case: dict, qtype: str) -> dict:
    torch.manual_seed(case['seed'])
    enc, q, dec = build(case, qtype)
    x = torch.rand(case['B'], 3, case['S'], case['S']) * 2 - 1
    enc.train(); q.train(); dec.train()

    init_sum = {}
    for pre, m in (('encoder.', enc), ('quantizer.', q), ('decoder.', dec)):
        for n, t in m.state_dict().items():
            init_sum[pre + n] = float(t.double().abs().sum())

    z = enc(x)
    noise = None
    if qtype == 'gumbel':
        st = torch.get_rng_state()
        quant, idx, ql = q(z)
        after = torch.get_rng_state()
        torch.set_rng_state(st)
        noise = torch.empty(case['B'], case['K'], z.shape[2], z.shape[3]).exponential_()
        torch.set_rng_state(after)
    else:
        quant, idx, ql = q(z)
    recon = dec(quant)
    l2 = torch.nn.functional.mse_loss(recon, x)
    (ql + l2).backward()

    out = {
        'z': z.detach().numpy(), 'idx': idx.numpy(), 'q_loss': np.float32(ql.item()),
        'quantized': quant.detach().numpy(), 'recon': recon.detach().numpy(), 'l2': np.float32(l2.item()),
        'grad_enc_conv_in': enc.conv_in.weight.grad.numpy(),
        'grad_dec_conv_out': dec.conv_out.weight.grad.numpy(),
        'grad_dec_conv_out_bias': dec.conv_out.bias.grad.numpy(),
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
    if qtype == 'ema':
        out['new_ema_count'] = q.ema_count.numpy()
        out['new_ema_weight'] = q.ema_weight.numpy()
        out['new_codebook'] = q.codebook.weight.detach().numpy()
    if qtype == 'standard' or qtype == 'entropy':
        out['grad_codebook'] = q.codebook.weight.grad.numpy()
    if noise is not None:
        out['exp_noise'] = noise.numpy()
        out['grad_x_to_logits'] = q.x_to_logits.weight.grad.numpy()
    return out


def vq_kernel_case(seed: int, N: int, K: int, D: int, init: str) -> dict:
    """Stand-alone quantizer goldens at the bench shapes (reference modules, forward only + EMA update)."""
    from vqvae.modules import vector_quantizers as vq
    torch.manual_seed(seed)
    q = vq.EMAVectorQuantizer(K, D, 0.25, 0.95, 1e-5)
    q.init_codebook()
    if init == 'normal':
        with torch.no_grad():
            q.codebook.weight.normal_()
            q.ema_weight.copy_(q.codebook.weight)
    b = N // 256
    z = torch.randn(b, D, 16, 16)
    q.train()
    quant, idx, loss = q(z)
    d_idx = idx.reshape(-1)
    return {'idx': d_idx.numpy().astype(np.int32), 'loss': np.float32(loss.item()),
            'new_ema_count': q.ema_count.numpy(),
            'codebook_rowsum': q.codebook.weight.detach().double().sum(1).numpy(),
            'quant_sum': np.float64(quant.double().sum().item())}


def main():
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    for name, case in CASES.items():
        for qtype in Q_PARAMS:
            res = run_case(name, case, qtype)
            path = os.path.join(OUT, f'{name}_{qtype}.npz')
            np.savez_compressed(path, **res)
            print(f'{path}: q_loss={res["q_loss"]:.8f} l2={res["l2"]:.8f} sum(idx)={int(res["idx"].sum())} '
                  f'z.sum={res["z"].sum():.6f}')
    for (N, K, init) in ((4096, 1024, 'uniform'), (4096, 1024, 'normal')):
        res = vq_kernel_case(77, N, K, 256, init)
        path = os.path.join(OUT, f'vqema_N{N}_K{K}_{init}.npz')
        np.savez_compressed(path, **res)
        print(path, 'loss', res['loss'], 'sum(idx)', int(res['idx'].sum()))


if __name__ == '__main__':
    main()
