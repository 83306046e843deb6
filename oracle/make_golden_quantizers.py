"""Generate tests/golden/quantizer_entropy_{softmax,argmax}.npz by EXECUTING THE REFERENCE'S OWN EntropyVectorQuantizer
(vqvae/modules/vector_quantizers.py:277-356) on seeded inputs, forward and backward (test infrastructure only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_quantizers.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = os.environ.get('VQ_REF_PATH', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def entropy_case(loss_type: str, seed=5, B=3, D=32, H=4, W=4, K=48):
    from vqvae.modules.vector_quantizers import EntropyVectorQuantizer
    torch.manual_seed(seed)
    q = EntropyVectorQuantizer(K, D, ent_loss_ratio=0.1, ent_temperature=0.01, ent_loss_type=loss_type, commitment_cost=0.25)
    with torch.no_grad():
        q.codebook.weight.copy_(torch.randn(K, D) * 0.03)          # distances / T of order one: a non-degenerate softmax
    z = (torch.randn(B, D, H, W) * 0.03).requires_grad_()
    gq = torch.randn(B, D, H, W)
    out, idx, loss = q(z)
    (loss * 1.5 + (out * gq).sum()).backward()
    return {'z': z.detach().numpy(), 'codebook': q.codebook.weight.detach().numpy(), 'g_q': gq.numpy(), 'q': out.detach().numpy(),
            'idx': idx.numpy(), 'loss': np.float32(loss.item()), 'dz': z.grad.numpy(), 'dcb': q.codebook.weight.grad.numpy()}


def main():
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    for t in ('softmax', 'argmax'):
        res = entropy_case(t)
        np.savez_compressed(os.path.join(OUT, f'quantizer_entropy_{t}.npz'), **res)
        print(t, res['loss'], np.abs(res['dz']).max(), np.abs(res['dcb']).max(), np.bincount(res['idx'].ravel(), minlength=48).max())


if __name__ == '__main__':
    main()
