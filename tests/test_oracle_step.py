"""CPU: the oracle's whole optimisation step (oracle/gan_oracle.py: schedules, forward, loss heads, adaptive weight, R1,
two AdamW with the reference's parameter grouping) against fixtures produced by EXECUTING the reference's own VQVAE class
(oracle/make_golden_step.py).  This is what pins the oracle for branches A / B / C of training_step incl. the optimizers.

Tolerances.  Scalars (losses, adaptive weight, R1, lr) and indices are held tight.  Weight CHANGES go through AdamW, whose
first steps are sign-like: an element whose gradient is below the fp32 re-association noise flips its update, so two correct
fp32 implementations that merely sum in a different order differ by ~2*sqrt(noise) in L2 on the deltas (1e-7 -> 6e-4 at the
stock eps = 1e-8; behind the ReLU / max-pool / leaky-ReLU stacks of LPIPS and the discriminator the gradient noise itself is
~1e-3, see oracle/step_cases.py).  The bars below are ~3x what this oracle measures against the reference here."""
import numpy as np
import pytest
import torch

from oracle import gan_oracle as G
from oracle import init_state as oinit
from oracle.step_cases import STEP_CASES, oracle_cfg_of
from tests import common as C

# aggregate / worst-tensor bars on the weight deltas, per step (measured here: see the module docstring)
BARS = {'mse': (2e-3, 1e-2), 'lpips': (5e-3, 3e-2), 'gan': (5e-2, 2e-1)}


def load_case(name):
    case = STEP_CASES[name]
    g = C.golden('step_' + name)
    crit = None if case['l_conf'] is None else ('gan' if case['l_conf']['adversarial_params'] is not None else 'lpips')
    sd = oinit.init_state(case['qtype'], case['K'], case['D'], case['ch'], case['nrb'], case['mult'], seed=case['seed'],
                          criterion=crit, image_size=case['S'])
    torch.manual_seed(case['seed'] + 1)
    xs = [torch.rand(case['B'], 3, case['S'], case['S']) for _ in range(case['steps'])]
    return case, g, crit, sd, xs


def gumbel_noise(case, step):
    """the Exp(1) tensor F.gumbel_softmax drew in the fixture run: first draw after manual_seed(7000 + step)"""
    if case['qtype'] != 'gumbel':
        return None
    torch.manual_seed(7000 + step)
    h = case['S'] // 2 ** len(case['mult'])
    return torch.empty(case['B'], case['K'], h, h).exponential_()


@pytest.mark.parametrize('name', list(STEP_CASES))
def test_seeded_init_matches_reference_instance(name):
    case, g, crit, sd, _ = load_case(name)
    ref = dict(zip(g['init_names'].tolist(), g['init_abs_sums'].tolist()))
    assert set(ref) == set(sd)
    for k, v in sd.items():
        assert abs(float(v.double().abs().sum()) - ref[k]) <= 1e-9 * max(1.0, ref[k]), k


@pytest.mark.parametrize('name', list(STEP_CASES))
def test_oracle_two_steps_match_reference(name):
    case, g, crit, sd, xs = load_case(name)
    torch.set_num_threads(8)
    init = {k: v.clone() for k, v in sd.items()}
    sd = oinit.make_leaf(sd, case['qtype'])
    opts = G.configure_optimizers(sd, case['t_conf'], gan=(crit == 'gan'))
    # parameter grouping = the reference's configure_optimizers incl. the relative-name collision (B2)
    for gi in range(2):
        assert sorted(opts[0].groups[gi][0]) == sorted(n for n in g[f'ae_group{gi}_names'].tolist() if sd[n].requires_grad)
        assert opts[0].groups[gi][1] == float(g[f'ae_group{gi}_wd'])
    names = g['log_names'].tolist()
    agg_bar, worst_bar = BARS[crit or 'mse']
    for i, x in enumerate(xs):
        log = G.train_step(sd, opts, x, oracle_cfg_of(case), case['l_conf'], case['t_conf'], 0, i, case['nb'],
                           exp_noise=gumbel_noise(case, i))
        ref = dict(zip(names, g['logs'][i].tolist()))
        assert np.array_equal(log['idx'].reshape(-1).numpy(), g[f'idx_{i}'].reshape(-1)), f'step {i} indices'
        for k in names:
            assert abs(float(log[k]) - ref[k]) <= 1e-4 * max(abs(ref[k]), 1e-2), (i, k, float(log[k]), ref[k])
        agg, worst, worst_dn = C.step_state_errors(lambda n: sd[n], g, i, init)
        assert agg <= agg_bar and worst[0] <= worst_bar and worst_dn[0] <= worst_bar, (i, agg, worst, worst_dn)
    if crit == 'gan':
        for n, fn, dn in zip(g['d_names'].tolist(), g['d_final_norms'].tolist(), g['d_delta_norms'].tolist()):
            assert abs(float(sd[n].double().norm()) - fn) <= 1e-5 * fn + 2e-2 * dn + 1e-12, n
            assert abs(float((sd[n].double() - init[n].double()).norm()) - dn) <= 2e-2 * dn + 1e-9, n
        for k in g.files:
            if k.startswith('d_delta/'):
                n = k[len('d_delta/'):]
                assert C.rel_err((sd[n].detach() - init[n])[:64], g[k]) < 2e-2, n
