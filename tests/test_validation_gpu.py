"""GPU: the hooks either side of training_step, against the CPU oracle on the seeded reference initialisation of the step
fixtures (oracle/step_cases.py): validation_step for each criterion type (vqvae/model.py:309-356), the two-stage inference API
(get_tokens / quantize / reconstruct / reconstruct_from_tokens, model.py:458-489), and FusedAdamW checkpointing in
torch.optim.AdamW's state format (resume = identical next step)."""
import pytest
import torch

from oracle import gan_oracle as G
from oracle import init_state as oinit
from oracle import vqvae_oracle as orc
from oracle.step_cases import STEP_CASES, oracle_cfg_of, q_conf_of
from tests import common as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def V():
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    pkg.lib.load()
    pkg.set_precision('strict')
    return pkg


def build(V, name, train=False):
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    case = STEP_CASES[name]
    crit = None if case['l_conf'] is None else ('gan' if case['l_conf']['adversarial_params'] is not None else 'lpips')
    sd = oinit.init_state(case['qtype'], case['K'], case['D'], case['ch'], case['nrb'], case['mult'], seed=case['seed'],
                          criterion=crit, image_size=case['S'])
    model = V.VQVAE(case['S'], dict(channels=case['ch'], num_res_blocks=case['nrb'], channel_multipliers=list(case['mult'])),
                    q_conf_of(case), case['l_conf'], dict(case['t_conf']), pretrained_lpips=False)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    model.training_augmentations = None
    tr = Trainer(max_epochs=1, num_training_batches=case['nb'])
    tr.attach(model)
    model.on_train_start()
    torch.manual_seed(case['seed'] + 1)
    x = torch.rand(case['B'], 3, case['S'], case['S'])
    return case, crit, sd, model.train() if train else model.eval(), tr, x


@pytest.mark.parametrize('name', ['mse_ema', 'mse_standard', 'lpips_ema', 'gan_hinge_adaptive_r1', 'gan_nonsat_fixed'])
def test_validation_step_matches_oracle(V, name):
    case, crit, sd, model, tr, x = build(V, name)
    loss = model.validation_step(x.cuda(), 0)
    # oracle, evaluation semantics: no EMA update, no adaptive weight, no R1 (loss.py:127,147)
    sdo = {k: v.clone() for k, v in sd.items()}
    images = orc.normalize_images(x)
    with torch.no_grad():
        out = orc.forward_vqvae(sdo, images, oracle_cfg_of(case), training=False)
        if crit is None:
            l2 = torch.nn.functional.mse_loss(out['recon'], images)
            ref = dict(loss=out['q_loss'] + l2, l1=0., l2=l2, p=0., g=0., d=0.)
        else:
            lo, l1, l2, p, g, _ = G.forward_autoencoder(sdo, case['l_conf'], out['q_loss'], images, out['recon'], 0, training=False)
            d = G.forward_discriminator(sdo, case['l_conf'], images, out['recon'], 0, 0, training=False)[1] if crit == 'gan' else 0.
            ref = dict(loss=lo, l1=l1, l2=l2, p=p, g=g, d=d)
    ref['q'] = out['q_loss']
    keys = dict(loss='validation/loss', l1='validation/l1_loss', l2='validation/l2_loss', q='validation/quant_loss',
                p='validation/perc_loss', g='validation/gen_loss', d='validation/disc_loss')
    assert abs(float(loss) - float(ref['loss'])) <= 1e-4 * abs(float(ref['loss']))
    for k, name_ in keys.items():
        got, want = float(torch.as_tensor(model.logged[name_]).reshape(-1)[0]), float(torch.as_tensor(ref[k]).reshape(-1)[0])
        assert abs(got - want) <= 1e-4 * max(abs(want), 0.1), (k, got, want)
    assert model.val_epoch_usage_count is not None and int(model.val_epoch_usage_count.sum()) == out['idx'].numel()
    assert torch.equal(torch.bincount(out['idx'].reshape(-1), minlength=case['K']), model.val_epoch_usage_count.cpu())
    model.on_validation_epoch_end()
    assert 'val_metrics/perplexity' in model.logged and model.val_epoch_usage_count is None


def test_inference_api_matches_oracle(V):
    case, crit, sd, model, tr, x = build(V, 'mse_standard')
    images = orc.normalize_images(x)
    with torch.no_grad():
        out = orc.forward_vqvae({k: v.clone() for k, v in sd.items()}, images, oracle_cfg_of(case), training=False)
    tokens = model.get_tokens(x.cuda())
    b, hw = tokens.shape
    assert torch.equal(tokens.cpu(), out['idx'])
    q = model.quantize(x.cuda())
    assert q.shape == (b, hw, case['D'])
    assert C.rel_err(q, sd['quantizer.codebook.weight'][out['idx']]) < 1e-6
    rec = model.reconstruct(x.cuda())
    ref = torch.clip(out['recon'] * 0.5 + 0.5, 0, 1)                              # base_autoencoder.py:52-61
    assert rec.shape == x.shape and C.rel_err(rec, ref) < 1e-4
    rec2 = model.reconstruct_from_tokens(tokens)                                  # defect B4 fixed: tokens -> [B,D,h,w] latent
    assert C.rel_err(rec2, ref) < 1e-4
    with pytest.raises(ValueError):
        model.reconstruct_from_tokens(tokens[:, :hw - 1])


def test_fused_adamw_state_dict_resumes_identically(V):
    case, crit, sd, model, tr, x = build(V, 'mse_standard', train=True)
    xg = x.cuda()
    for i in range(2):
        tr.run_step(xg, i)
    opt_state = tr.optimizers[0].state_dict()
    model_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    # torch.optim.AdamW's layout: per-parameter step / exp_avg / exp_avg_sq, integer parameter ids in the groups
    assert set(opt_state) == {'state', 'param_groups'} and len(opt_state['param_groups']) == 2
    n_params = sum(len(g['params']) for g in opt_state['param_groups'])
    assert len(opt_state['state']) == n_params
    st0 = opt_state['state'][0]
    assert set(st0) == {'step', 'exp_avg', 'exp_avg_sq'} and float(st0['step']) == 2.0
    ref_opt = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))])            # a stock optimizer accepts the same group keys
    assert {'lr', 'betas', 'eps', 'weight_decay'} <= set(opt_state['param_groups'][0]) and 'lr' in ref_opt.state_dict()['param_groups'][0]
    tr.run_step(xg, 2)
    want = {k: v.detach().clone() for k, v in model.state_dict().items()}
    # fresh model + optimizer, state restored, same third step
    case, crit, sd, model2, tr2, _ = build(V, 'mse_standard', train=True)
    model2.load_state_dict(model_state, strict=True)
    V.ops.bump_weights_epoch()
    tr2.optimizers[0].load_state_dict(opt_state)
    assert tr2.optimizers[0].step_count == 2
    tr2.run_step(xg, 2)
    got = model2.state_dict()
    for k in want:                         # (weight gradients are combined with floating-point atomics: not bit-reproducible)
        assert C.rel_err(got[k], want[k]) < 1e-5, k


def test_trainer_validate_and_test_loops(V):
    """Trainer.validate / Trainer.test (pl.Trainer.validate / .test around the module's hooks) on HOST batches -- (images, labels)
    tuples as the reference's loaders yield -- log what the hooks log when called by hand on device batches."""
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    case, crit, sd, model, tr, x = build(V, 'lpips_ema')
    torch.manual_seed(21)
    host = [(torch.rand(case['B'], 3, case['S'], case['S']), torch.tensor(i)) for i in range(3)]
    # by hand
    for i, (xb, _) in enumerate(host):
        model.validation_step(xb.cuda(), i)
    want_val = {k: float(torch.as_tensor(v).reshape(-1)[0]) for k, v in model.logged.items() if k.startswith('validation/')}
    model.on_validation_epoch_end()
    want_ppl = float(model.logged['val_metrics/perplexity'])
    model.on_test_epoch_start()
    for xb, _ in host:
        model.test_step(xb.cuda(), 0)
    model.on_test_epoch_end()
    want_test = {k: float(model.logged[k]) for k in ('mse', 'psnr', 'ssim', 'perplexity', 'used_codebook')}
    # through the loops
    model.train()
    model.logged.clear()
    logged = Trainer().validate(model, host)
    assert model.training                                          # the loop restores the mode it found
    for k, v in want_val.items():
        assert abs(float(torch.as_tensor(logged[k]).reshape(-1)[0]) - v) <= 1e-5 * max(abs(v), 0.1), k
    assert abs(float(logged['val_metrics/perplexity']) - want_ppl) <= 1e-5 * want_ppl
    logged = Trainer().test(model, host)
    for k, v in want_test.items():
        assert abs(float(logged[k]) - v) <= 1e-5 * max(abs(v), 0.1), k
