"""Case table of the two-step optimisation fixtures (tests/golden/step_*.npz)  --  TEST INFRASTRUCTURE ONLY.
Shared by oracle/make_golden_step.py (which executes the reference's VQVAE class on these), tests/test_oracle_step.py (CPU
oracle vs the fixtures) and tests/test_train_step_gpu.py (CUDA path vs the fixtures)."""
from __future__ import annotations

Q_PARAMS = {
    'standard': dict(commitment_cost=0.25),
    'ema': dict(commitment_cost=0.25, decay=0.95, epsilon=1e-5),
    'entropy': dict(ent_loss_ratio=0.1, ent_temperature=0.01, ent_loss_type='softmax', commitment_cost=0.25),
    # schedules active: cosine kl warm-up over 1 epoch (nb steps) and temperature decay 1.0 -> 0.5 (model.py:190-200,219-225)
    'gumbel': dict(straight_through=False, temp=1.0, kl_cost=0.00859375, kl_warmup_epochs=1, temp_decay_epochs=1, temp_final=0.5),
}
T_CONF = dict(lr=1e-3, betas=[0.5, 0.9], eps=1e-8, weight_decay=1e-2, warmup_epochs=None, decay_epochs=None)
T_CONF_SCHED = dict(T_CONF, warmup_epochs=1, decay_epochs=3)          # linear warm-up then cosine (model.py:170-175)
# AdamW's first steps are sign-like (dw = lr * g / (|g| + eps)): an element whose gradient is smaller than the fp32
# re-association noise flips its whole update.  Behind ReLU / max-pool / leaky-ReLU stacks at random init that noise is ~1e-3
# of the gradient (the reference's own fp32 LPIPS gradient is 0.7 % from its fp64 value), which would put ~10 % L2 error on the
# weight DELTAS of any two correct fp32 implementations.  The LPIPS / GAN cases therefore use eps = 1e-4 (the order of one
# gradient element), which keeps the update Lipschitz in the gradient; the MSE cases keep the usual 1e-8.
T_CONF_SMOOTH = dict(T_CONF, eps=1e-5)
T_CONF_LOSS_HEADS = dict(T_CONF, lr=1e-4, eps=1e-4)     # lr as in example_confs/*.yaml: keeps step 1 from amplifying step 0's noise
TINY = dict(S=16, B=4, ch=32, nrb=1, mult=(1, 2), K=64, D=32, nb=4, steps=2)
SMALL64 = dict(S=64, B=4, ch=32, nrb=1, mult=(1, 2), K=64, D=32, nb=4, steps=2)


def _adv(loss_type, adaptive, r1, start_epoch=0):
    return dict(start_epoch=start_epoch, loss_type=loss_type, g_weight=0.1, use_adaptive=adaptive, r1_reg_weight=r1, r1_reg_every=16)


def _l(adv):
    return dict(l1_weight=0.8, l2_weight=0.2, perc_weight=1.0, adversarial_params=adv)


STEP_CASES = {
    # branch C (plain MSE), every quantizer type
    'mse_standard': dict(TINY, qtype='standard', seed=101, l_conf=None, t_conf=T_CONF_SMOOTH),
    'mse_ema': dict(TINY, qtype='ema', seed=102, l_conf=None, t_conf=T_CONF_SCHED),
    'mse_entropy': dict(TINY, qtype='entropy', seed=103, l_conf=None, t_conf=T_CONF),
    'mse_gumbel': dict(TINY, qtype='gumbel', seed=104, l_conf=None, t_conf=T_CONF_SMOOTH),
    # branch B (LPIPS-AlexNet, no discriminator)
    'lpips_ema': dict(SMALL64, qtype='ema', seed=105, l_conf=_l(None), t_conf=T_CONF_LOSS_HEADS),
    # branch A (VQGAN): hinge + adaptive weight + R1 (step 0); non-saturating, fixed weight, no R1; discriminator not started yet
    'gan_hinge_adaptive_r1': dict(SMALL64, qtype='ema', seed=106, l_conf=_l(_adv('hinge', True, 10.0)), t_conf=T_CONF_LOSS_HEADS),
    'gan_nonsat_fixed': dict(SMALL64, qtype='standard', seed=107, l_conf=_l(_adv('non-saturating', False, None)), t_conf=T_CONF_LOSS_HEADS),
    'gan_not_started': dict(SMALL64, qtype='gumbel', seed=108, l_conf=_l(_adv('hinge', True, 10.0, start_epoch=1)), t_conf=T_CONF_LOSS_HEADS),
}


def q_conf_of(case: dict) -> dict:
    return dict(num_embeddings=case['K'], embedding_dim=case['D'], type=case['qtype'], params=dict(Q_PARAMS[case['qtype']]),
                reinit_every_n_epochs=None)


def oracle_cfg_of(case: dict) -> dict:
    return {'num_res_blocks': case['nrb'], 'channel_multipliers': case['mult'],
            'quantizer': dict(Q_PARAMS[case['qtype']], type=case['qtype'])}
