"""GPU: the bf16 / tcgen05 'fast' precision mode end to end against the reference fixtures.  The reference itself trains
with precision='16-mixed' (vqvae/train.py:129); bf16 activations cannot meet the fp32 1e-4 bar, so this test states the
tolerance of the fast mode explicitly: losses within 2 %, tensors within a few % relative L2, gradient norms within 10 %."""
import pytest
import torch

from tests import common as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def V():
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    pkg.lib.load()
    if not pkg.lib.load().vqb_device_supports_tcgen05():
        pytest.skip('needs sm_100')
    pkg.set_precision('fast')
    yield pkg
    pkg.set_precision('strict')


@pytest.mark.parametrize('qtype', ['standard', 'ema', 'entropy'])
def test_fast_mode_train_step_close_to_reference(V, qtype):
    g = C.golden(f'cfg1_{qtype}')
    sd, x = C.seeded_inputs('cfg1', qtype)
    c = C.CASES['cfg1']
    qp = {k: v for k, v in C.Q_PARAMS[qtype].items() if k != 'type'}
    model = V.VQVAE(c['S'], dict(channels=c['ch'], num_res_blocks=c['nrb'], channel_multipliers=list(c['mult'])),
                    dict(num_embeddings=c['K'], embedding_dim=c['D'], type=qtype, params=qp, reinit_every_n_epochs=None),
                    None, dict(lr=1e-4, betas=[0.0, 0.99], eps=1e-8, weight_decay=1e-4, warmup_epochs=None, decay_epochs=None))
    model.load_state_dict(sd)
    model = model.cuda().train()
    xg = x.cuda().contiguous(memory_format=torch.channels_last)
    recon, q_loss, idx = model(xg)
    l2 = model.criterion(recon, xg)
    (q_loss + l2).backward()
    z = model.encoder(xg).detach()
    assert z.dtype == torch.float32 and recon.dtype == torch.float32
    assert C.rel_err(z, g['z']) < 3e-2
    assert abs(float(l2.detach()) - float(g['l2'])) <= 2e-2 * float(g['l2'])
    assert abs(float(q_loss.detach()) - float(g['q_loss'])) <= 5e-2 * abs(float(g['q_loss'])) + 1e-4
    ref_norm = dict(zip(g['grad_names'].tolist(), g['grad_norms'].tolist()))
    worst = 0.0
    for n, p in model.named_parameters():
        if p.grad is not None and n in ref_norm and ref_norm[n] > 1e-6:
            worst = max(worst, abs(float(p.grad.double().norm()) - ref_norm[n]) / ref_norm[n])
    assert worst < 0.10, worst
