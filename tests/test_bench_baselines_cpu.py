"""CPU: the optimizer adapter of bench.py's secondary baseline (`torch_eager_gpu_baseline`: the oracle port of the reference modules
run as stock torch eager ops) performs the reference's update -- torch.optim.AdamW over the reference's parameter groups equals
the oracle's own AdamW (which tests/test_oracle_step.py pins against fixtures made by executing the reference's VQVAE class)."""
import pytest
import torch

import bench
from oracle import gan_oracle as G
from oracle import init_state as oinit
from oracle.step_cases import STEP_CASES, oracle_cfg_of


@pytest.mark.parametrize('name', ['mse_ema', 'mse_standard'])
def test_torch_adamw_adapter_equals_oracle_adamw(name):
    case = STEP_CASES[name]
    torch.set_num_threads(4)
    states = []
    for use_torch in (False, True):
        sd = oinit.init_state(case['qtype'], case['K'], case['D'], case['ch'], case['nrb'], case['mult'], seed=case['seed'],
                              criterion=None, image_size=case['S'])
        sd = oinit.make_leaf(sd, case['qtype'])
        init = {k: v.detach().clone() for k, v in sd.items()}
        opts = G.configure_optimizers(sd, case['t_conf'], gan=False)
        if use_torch:
            opts = [bench._TorchAdamW(sd, o) for o in opts]
        torch.manual_seed(case['seed'] + 1)
        for i in range(3):
            x = torch.rand(case['B'], 3, case['S'], case['S'])
            G.train_step(sd, opts, x, oracle_cfg_of(case), None, case['t_conf'], 0, i, case['nb'])
        states.append({k: v.detach().clone() for k, v in sd.items()})
    a, b = states
    num = den = 0.0
    for k in a:
        if a[k].dtype.is_floating_point and not k.startswith('quantizer.ema_'):
            num += float((a[k].double() - b[k].double()).pow(2).sum())
            den += float((a[k].double() - init[k].double()).pow(2).sum())
    assert den > 0
    # same arithmetic up to the operation order inside torch's foreach kernels (measured 5e-8 .. 1e-6 of the weight CHANGES)
    assert (num / den) ** 0.5 < 1e-4, (num / den) ** 0.5
