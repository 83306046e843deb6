"""Shared helpers for the tests: golden loading, case definitions, tie-aware index comparison."""
import os

import numpy as np
import torch

from oracle import init_state as oinit
from oracle import vqvae_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

CASES = {
    'cfg1': dict(S=64, B=8, ch=128, nrb=2, mult=(1, 2, 2, 4), K=256, D=256, seed=1234),
    'tiny': dict(S=16, B=4, ch=32, nrb=1, mult=(1, 2), K=64, D=32, seed=4321),
}
Q_PARAMS = {
    'standard': dict(type='standard', commitment_cost=0.25),
    'ema': dict(type='ema', commitment_cost=0.25, decay=0.95, epsilon=1e-5),
    'entropy': dict(type='entropy', ent_loss_ratio=0.1, ent_temperature=0.01, ent_loss_type='softmax',
                    commitment_cost=0.25),
    'gumbel': dict(type='gumbel', straight_through=False, temp=1.0, kl_cost=0.00859375),
}


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def seeded_inputs(case_name, qtype):
    """Regenerate (state dict, images in [-1,1]) exactly as oracle/make_golden.py did."""
    c = CASES[case_name]
    sd = oinit.init_state(qtype, c['K'], c['D'], c['ch'], c['nrb'], c['mult'], seed=c['seed'])
    x = torch.rand(c['B'], 3, c['S'], c['S']) * 2 - 1
    return sd, x


# BASELINE configs[1]'s architecture and image size at batch 2 (oracle/make_golden_256.py: tests/golden/cfg2_256_ema.npz)
CASES['cfg2_256'] = dict(S=256, B=2, ch=128, nrb=2, mult=(1, 2, 2, 4), K=1024, D=256, seed=2468)


def seeded_inputs_256():
    """(state dict, images in [-1,1]) exactly as oracle/make_golden_256.py made them: the cfg1 recipe, then a tie-free codebook"""
    sd, x = seeded_inputs('cfg2_256', 'ema')
    c = CASES['cfg2_256']
    torch.manual_seed(c['seed'] + 1)
    sd['quantizer.codebook.weight'] = torch.randn(c['K'], c['D']) * 0.05
    sd['quantizer.ema_weight'] = sd['quantizer.codebook.weight'].clone()
    return sd, x


def oracle_cfg(case_name, qtype):
    c = CASES[case_name]
    return {'num_res_blocks': c['nrb'], 'channel_multipliers': c['mult'], 'quantizer': dict(Q_PARAMS[qtype])}


def rel_err(a, b):
    a = torch.as_tensor(a).detach().double().flatten().cpu()
    b = torch.as_tensor(b).detach().double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel(a, b, floor=None):
    """max |a-b| / max(|b|, floor); floor defaults to 1e-3 * max|b| (elementwise rel. error is meaningless
    near zero crossings)."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    if floor is None:
        floor = 1e-3 * float(b.abs().max()) + 1e-30
    return float(((a - b).abs() / torch.clamp(b.abs(), min=floor)).max())


def tie_aware_index_check(idx, ref_idx, flat, codebook, order='standard', ulps=4):
    """Indices must be bit-exact except where the reference's own fp32 distances are tied/near-tied
    (SURVEY.md section 7 'argmin tie fragility').  Returns (n_exact, n_tie_class, n_bad)."""
    idx = torch.as_tensor(idx).reshape(-1).long().cpu()
    ref_idx = torch.as_tensor(ref_idx).reshape(-1).long().cpu()
    mism = (idx != ref_idx).nonzero().flatten()
    n_bad = 0
    if mism.numel():
        d = orc.l2_distances(flat[mism].cpu().float(), codebook.cpu().float(), order)
        da = d.gather(1, idx[mism, None]).squeeze(1)
        db = d.gather(1, ref_idx[mism, None]).squeeze(1)
        scale = torch.maximum(da.abs(), db.abs()).clamp_min(1e-30)
        tol = ulps * torch.finfo(torch.float32).eps * torch.maximum(scale, (flat[mism].cpu().float() ** 2).sum(1))
        n_bad = int(((da - db).abs() > tol).sum())
    return int(idx.numel() - mism.numel()), int(mism.numel() - n_bad), n_bad


# ---- two-step optimisation fixtures (tests/golden/step_*.npz, oracle/make_golden_step.py) -------------------------------
# a conv bias that feeds a GroupNorm whose groups hold ONE channel (ch = 32) is cancelled by the mean subtraction: its
# gradient is mathematically zero, numerically ~1e-8 noise, and AdamW turns that noise into +-lr steps
DEGENERATE = ('decoder.blocks.3.conv.bias',)
HEAD = 512


def step_state_errors(get, g, i, init):
    """Compare the state after step `i` with the fixture.  get(name) -> current tensor, init[name] -> initial tensor.
    Returns (aggregate rel. L2 error of the weight CHANGES over the stored leading elements of every tensor,
             worst per-tensor error of the same [only tensors that moved], worst rel. error of the per-tensor change norms)."""
    names = g['state_names'].tolist()
    dn_ref = dict(zip(names, g[f'dnorm_{i}'].tolist()))
    num = den = 0.0
    worst, worst_dn = (0.0, None), (0.0, None)
    for n in names:
        if n in DEGENERATE:
            continue
        cur = torch.as_tensor(get(n)).detach().double().cpu().reshape(-1)
        ini = torch.as_tensor(init[n]).detach().double().cpu().reshape(-1)
        d = (cur - ini)[:HEAD]
        dref = torch.from_numpy(g[f'w{i}/{n}']).double() - ini[:HEAD]
        num += float((d - dref).pow(2).sum())
        den += float(dref.pow(2).sum())
        if float(dref.norm()) > 0:
            e = float((d - dref).norm() / dref.norm())
            if e > worst[0]:
                worst = (e, n)
        if dn_ref[n] > 0:
            e = abs(float((cur - ini).norm()) - dn_ref[n]) / dn_ref[n]
            if e > worst_dn[0]:
                worst_dn = (e, n)
        else:
            assert float((cur - ini).norm()) == 0.0, f'{n} must not move (the reference leaves it untouched)'
    return (num / max(den, 1e-300)) ** 0.5, worst, worst_dn
