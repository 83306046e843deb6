"""Generate tests/golden/step_*.npz by EXECUTING THE REFERENCE'S OWN `vqvae.model.VQVAE` for two optimisation steps
(test infrastructure only; run in the build container, the reference cannot travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_step.py

The reference class is imported unmodified through oracle/ref_harness.py (stand-ins for the absent third-party packages).
Per case: seed -> VQVAE(...) in the reference's construction order -> configure_optimizers() (the reference's grouping,
defect B2 included) -> on_train_start -> for each batch: on_train_batch_start, then
  * branch A (VQGAN): the reference's training_step VERBATIM (manual optimisation, two AdamW, R1 on step 0);
  * branches B / C: the reference's forward + criterion and what Lightning's automatic optimisation does around
    training_step (zero_grad, backward, step) -- training_step itself cannot be used there because it returns an unbound
    `loss` (defect B1, model.py:295).
Everything stored is an OUTPUT of reference code.  Inputs are regenerated from seeds by the consumers: the initial state by
oracle/init_state.init_state (checked here tensor by tensor against the reference instance), the image batches from
`seed + 1`, the Gumbel noise from `7000 + step` (F.gumbel_softmax's exponential_() is the first draw after the seed).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import init_state as oinit            # noqa: E402
from oracle import ref_harness as H               # noqa: E402
from oracle.step_cases import STEP_CASES, q_conf_of          # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
HEAD = 512


def flt(v) -> float:
    return float(v.detach()) if torch.is_tensor(v) else float(v)


def run_case(name: str, case: dict) -> dict:
    VQVAE = H.reference_vqvae_class()
    ae_conf = dict(channels=case['ch'], num_res_blocks=case['nrb'], channel_multipliers=list(case['mult']))
    torch.manual_seed(case['seed'])
    model = VQVAE(case['S'], ae_conf, q_conf_of(case), case['l_conf'], dict(case['t_conf']))
    model.train()

    # the oracle's seeded construction must equal the reference instance bit for bit
    crit = None if case['l_conf'] is None else ('gan' if case['l_conf']['adversarial_params'] is not None else 'lpips')
    sd0 = oinit.init_state(case['qtype'], case['K'], case['D'], case['ch'], case['nrb'], case['mult'], seed=case['seed'],
                           criterion=crit, image_size=case['S'])
    ref0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert set(sd0) == set(ref0), (sorted(set(sd0) ^ set(ref0))[:10])
    assert all(torch.equal(sd0[k], ref0[k]) for k in ref0)

    nb = case['nb']
    trainer = H.make_trainer(nb)
    model.trainer = trainer
    opt = model.configure_optimizers()
    trainer.optimizers = list(opt[0]) if isinstance(opt, tuple) else [opt]
    pname = {id(p): n for n, p in model.named_parameters()}
    out = {}
    for gi, g in enumerate(trainer.optimizers[0].param_groups):
        out[f'ae_group{gi}_names'] = np.array([pname[id(p)] for p in g['params']])
        out[f'ae_group{gi}_wd'] = np.float64(g['weight_decay'])
    model.on_train_start()

    # gradients as the optimizers see them (captured right before every optimizer.step())
    grads = {}

    def capture(tag):
        def hook(optimizer, args, kwargs):
            for grp in optimizer.param_groups:
                for p in grp['params']:
                    if p.grad is not None:
                        grads[(tag, step_box[0], pname[id(p)])] = p.grad.detach().clone()
        return hook

    step_box = [0]
    for tag, o in zip(('ae', 'd'), trainer.optimizers):
        o.register_step_pre_hook(capture(tag))

    torch.manual_seed(case['seed'] + 1)
    xs = [torch.rand(case['B'], 3, case['S'], case['S']) for _ in range(case['steps'])]
    captured = {}
    model.quantizer.register_forward_hook(lambda m, i, o: captured.__setitem__('idx', o[1].detach().clone()))
    gan = crit == 'gan'
    logs = []
    for i, x in enumerate(xs):
        step_box[0] = i
        torch.manual_seed(7000 + i)
        model.on_train_batch_start(x, i)
        lr = trainer.optimizers[0].param_groups[0]['lr']
        if gan:
            model.training_step(x, i)                           # model.py:232-295 verbatim (two optimizers inside)
            L = model.logged
            row = [L['train/loss'], L['train/l1_loss'], L['train/l2_loss'], L['train/quant_loss'], L['train/perc_loss'],
                   L['train/gen_loss'], L['train/disc_loss'], flt(L['g_weight']), flt(L['r1_penalty']), lr]
        else:
            o = trainer.optimizers[0]
            o.zero_grad()
            images = model.preprocess_batch(x, training=True)
            x_recon, q_loss, _ = model.forward(images)
            if crit == 'lpips':
                loss, l1, l2, p = model.criterion(q_loss, images, x_recon)             # model.py:266-269
            else:
                l2 = model.criterion(x_recon, images)                                  # model.py:271-275
                loss, l1, p = q_loss + l2, torch.zeros(1), torch.zeros(1)
            loss.backward()
            o.step()
            row = [flt(loss), flt(l1), flt(l2), flt(q_loss), flt(p), 0., 0., 0., 0., lr]
        torch.use_deterministic_algorithms(False)               # training_step leaves it switched on (model.py:291)
        logs.append([float(v) for v in row])
        out[f'idx_{i}'] = captured['idx'].numpy().astype(np.int32)
        # state after this step: the leading elements of every autoencoder / quantizer tensor and the norm of its change
        cur = {k: v.detach() for k, v in model.state_dict().items() if not k.startswith('criterion.')}
        for k, v in cur.items():
            out[f'w{i}/{k}'] = v.reshape(-1)[:HEAD].clone().numpy()
        out[f'dnorm_{i}'] = np.array([float((v.double() - ref0[k].double()).norm()) for k, v in cur.items()])
        if case['qtype'] == 'gumbel':
            out[f'gumbel_consts_{i}'] = np.array(model.quantizer.get_consts(), dtype=np.float64)
    out['log_names'] = np.array(['loss', 'l1', 'l2', 'q', 'p', 'g', 'd', 'g_weight', 'r1', 'lr'])
    out['logs'] = np.array(logs, dtype=np.float64)

    for i in range(case['steps']):
        for tag in ('ae', 'd'):
            ks = [k for k in grads if k[0] == tag and k[1] == i]
            if ks:
                out[f'grad_names_{tag}_{i}'] = np.array([k[2] for k in ks])
                out[f'grad_norms_{tag}_{i}'] = np.array([float(grads[k].double().norm()) for k in ks])
    for n in ('encoder.conv_in.weight', 'decoder.conv_out.weight', 'decoder.conv_in.bias', 'encoder.norm.weight'):
        for i in range(case['steps']):
            if ('ae', i, n) in grads:
                out[f'grad_{i}/{n}'] = grads[('ae', i, n)].numpy()
    final = {k: v.detach() for k, v in model.state_dict().items()}
    out['state_names'] = np.array([k for k in final if not k.startswith('criterion.')])
    out['init_names'] = np.array(list(ref0.keys()))
    out['init_abs_sums'] = np.array([float(ref0[k].double().abs().sum()) for k in ref0], dtype=np.float64)
    if gan:
        dn = [k for k in final if k.startswith('criterion.discriminator.') and not k.endswith('resample_filter')]
        out['d_names'] = np.array(dn)
        out['d_final_norms'] = np.array([float(final[k].double().norm()) for k in dn])
        out['d_delta_norms'] = np.array([float((final[k].double() - ref0[k].double()).norm()) for k in dn])
        for k in ('criterion.discriminator.b4.out.weight', 'criterion.discriminator.b4.fc.bias',
                  f'criterion.discriminator.b{case["S"]}.fromrgb.weight', 'criterion.discriminator.b8.skip.weight'):
            out['d_delta/' + k] = (final[k] - ref0[k]).numpy()[:64]
        # the frozen LPIPS trunk must not move
        assert all(torch.equal(final[k], ref0[k]) for k in final if k.startswith('criterion.perceptual_loss.'))
    return out


def main():
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, case in STEP_CASES.items():
        if only and name not in only:
            continue
        res = run_case(name, case)
        path = os.path.join(OUT, f'step_{name}.npz')
        np.savez_compressed(path, **res)
        print(f'{path}: {os.path.getsize(path) / 1e3:.0f} KB')
        for row in res['logs']:
            print('   ', ' '.join(f'{n}={v:.6g}' for n, v in zip(res['log_names'], row)))


if __name__ == '__main__':
    main()
