"""GPU: Trainer(cuda_graph=True) -- one captured CUDA graph per step variant, replayed -- must compute what the eager step
computes: same weights / EMA state / optimizer state after the same number of steps, under a learning-rate schedule (AdamW's
scalars in device memory), the Gumbel temperature / KL schedules, and the VQGAN step with its R1 variant."""
import pytest
import torch

from oracle import init_state as oinit
from oracle.step_cases import STEP_CASES, q_conf_of
from tests import common as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def V():
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    pkg.lib.load()
    yield pkg
    pkg.set_precision('strict')


def run(V, name, graph, steps, mode, sync=True):
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    V.set_precision(mode)
    case = dict(STEP_CASES[name])
    if name.startswith('gan'):
        case['l_conf'] = dict(case['l_conf'], adversarial_params=dict(case['l_conf']['adversarial_params'], r1_reg_every=4))
    crit = None if case['l_conf'] is None else ('gan' if case['l_conf']['adversarial_params'] is not None else 'lpips')
    sd = oinit.init_state(case['qtype'], case['K'], case['D'], case['ch'], case['nrb'], case['mult'], seed=case['seed'],
                          criterion=crit, image_size=case['S'])
    model = V.VQVAE(case['S'], dict(channels=case['ch'], num_res_blocks=case['nrb'], channel_multipliers=list(case['mult'])),
                    q_conf_of(case), case['l_conf'], dict(case['t_conf']), pretrained_lpips=False)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    model.training_augmentations = None
    tr = Trainer(max_epochs=1, num_training_batches=steps, cuda_graph=graph, graph_warmup=2)
    tr.attach(model)
    model.on_train_start()
    torch.manual_seed(case['seed'] + 1)
    xs = [torch.rand(case['B'], 3, case['S'], case['S']).cuda() for _ in range(2)]
    losses = []
    for i in range(steps):
        if case['qtype'] == 'gumbel':
            torch.manual_seed(500 + i)               # the Gumbel noise comes from torch's CUDA generator in both runs
            torch.cuda.manual_seed(500 + i)
        loss = tr.run_step(xs[i % 2], i)
        if sync:
            losses.append(float(loss))              # a device -> host read: the host never runs ahead of the device
    return {k: v.detach().clone() for k, v in model.state_dict().items()}, losses, tr


@pytest.mark.parametrize('name,mode', [('mse_ema', 'strict'), ('mse_ema', 'fast'), ('mse_standard', 'strict'), ('mse_entropy', 'strict'),
                                       ('mse_gumbel', 'strict'),
                                       ('gan_hinge_adaptive_r1', 'strict'), ('gan_nonsat_fixed', 'fast')])
def test_graph_replay_equals_eager(V, name, mode):
    if mode == 'fast' and not V.lib.load().vqb_device_supports_tcgen05():
        pytest.skip('needs sm_100')
    steps = 10 if name.startswith('gan') else 7          # gan: R1 on steps 0, 4, 8 -> the R1 variant is captured on step 8
    eager, le, _ = run(V, name, False, steps, mode)
    graph, lg, tr = run(V, name, True, steps, mode)
    assert any(st['graph'] is not None for st in tr._graphs.values()), 'no graph was captured'
    for a, b in zip(le, lg):            # fast mode: bf16 activations amplify the run-to-run noise of the atomics (two EAGER runs differ alike)
        assert abs(a - b) <= (2e-3 if mode == 'strict' else 2e-2) * max(abs(a), 0.1), (le, lg)
    # (two runs differ by the floating-point atomics of the weight-gradient combine, amplified by AdamW's sign-like steps: the
    # bars are those of two eager runs, see tests/test_data_parallel_gpu.py)
    num = den = 0.0
    for k in eager:
        if k in C.DEGENERATE or not eager[k].dtype.is_floating_point:
            continue
        num += float((graph[k].double() - eager[k].double()).pow(2).sum()); den += float(eager[k].double().pow(2).sum())
    # strict: 1e-4 .. 2.5e-4 measured between two runs of the 10-step GAN case (state dominated by the frozen trunk + initial weights)
    assert (num / den) ** 0.5 < (1e-3 if mode == 'strict' else 5e-2), (num / den) ** 0.5


@pytest.mark.parametrize('name', ['mse_ema', 'mse_gumbel'])
def test_graph_replay_with_the_host_running_ahead(V, name):
    """No synchronisation between steps: the host queues every replay while the device is still on the first ones.  The per-step
    scalars (learning-rate schedule + Adam bias corrections; Gumbel temperature / KL weight) must still be the ones of THEIR
    step -- they reach the device through ops.StepScalars' ring of pinned slots, not through one pinned buffer that the host
    would have re-written before the device read it."""
    steps = 12
    synced, _, _ = run(V, name, True, steps, 'strict', sync=True)
    ahead, _, tr = run(V, name, True, steps, 'strict', sync=False)
    torch.cuda.synchronize()
    assert any(st['graph'] is not None for st in tr._graphs.values()), 'no graph was captured'
    num = den = 0.0
    for k in synced:
        if k in C.DEGENERATE or not synced[k].dtype.is_floating_point:
            continue
        num += float((ahead[k].double() - synced[k].double()).pow(2).sum()); den += float(synced[k].double().pow(2).sum())
    # two runs of the same graphs differ by the floating-point atomics of the weight-gradient combine (~1e-4 after a few AdamW
    # steps); a step that read a LATER step's learning rate / bias corrections moves the state by ~1e-2
    assert (num / den) ** 0.5 < 3e-3, (num / den) ** 0.5


def test_fit_loop_with_graphs(V):
    """Trainer.fit (which keeps the previous step's loss while the next step runs) captures and replays graphs"""
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    V.set_precision('strict')
    case = STEP_CASES['mse_ema']
    sd = oinit.init_state(case['qtype'], case['K'], case['D'], case['ch'], case['nrb'], case['mult'], seed=case['seed'],
                          criterion=None, image_size=case['S'])
    model = V.VQVAE(case['S'], dict(channels=case['ch'], num_res_blocks=case['nrb'], channel_multipliers=list(case['mult'])),
                    q_conf_of(case), None, dict(case['t_conf']), pretrained_lpips=False)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    model.training_augmentations = None
    torch.manual_seed(3)
    batches = [torch.rand(case['B'], 3, case['S'], case['S']).cuda() for _ in range(6)]
    tr = Trainer(max_epochs=2, cuda_graph=True, graph_warmup=2)
    loss = tr.fit(model, batches)
    assert any(st['graph'] is not None for st in tr._graphs.values()), 'no graph was captured'
    assert loss.grad_fn is None and torch.isfinite(loss).all()
