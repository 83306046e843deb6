"""CPU: pin the oracle (oracle/vqvae_oracle.py) against fixtures produced by the reference's own modules."""
import os

import numpy as np
import pytest
import torch

from oracle import init_state as oinit
from oracle import vqvae_oracle as orc
from tests import common as C


@pytest.mark.parametrize('case', ['tiny', 'cfg1'])
@pytest.mark.parametrize('qtype', ['standard', 'ema', 'entropy', 'gumbel'])
def test_oracle_matches_reference_fixture(case, qtype):
    g = C.golden(f'{case}_{qtype}')
    sd, x = C.seeded_inputs(case, qtype)
    # initial weights identical to the reference's constructors (same RNG stream)
    ref_init = dict(zip(g['init_names'].tolist(), g['init_abs_sums'].tolist()))
    assert set(ref_init) == set(sd)
    for n, t in sd.items():
        assert abs(float(t.double().abs().sum()) - ref_init[n]) <= 1e-9 * max(1.0, ref_init[n]), n
    sd = oinit.make_leaf(sd, qtype)
    noise = torch.from_numpy(g['exp_noise']) if qtype == 'gumbel' else None
    out = orc.train_step_mse(sd, x, C.oracle_cfg(case, qtype), exp_noise=noise)

    assert np.array_equal(out['idx'].numpy(), g['idx']), 'codebook indices must be bit-exact'
    assert C.rel_err(out['z'], g['z']) < 1e-5
    assert C.rel_err(out['recon'], g['recon']) < 2e-5
    assert abs(out['q_loss'].item() - float(g['q_loss'])) <= 2e-6 * max(1.0, abs(float(g['q_loss'])))
    assert abs(out['l2_loss'].item() - float(g['l2'])) <= 1e-6
    assert C.rel_err(sd['encoder.conv_in.weight'].grad, g['grad_enc_conv_in']) < 1e-4
    assert C.rel_err(sd['decoder.conv_out.weight'].grad, g['grad_dec_conv_out']) < 1e-4
    ref_norm = dict(zip(g['grad_names'].tolist(), g['grad_norms'].tolist()))
    for n, t in sd.items():
        if t.grad is not None and n in ref_norm:
            assert abs(float(t.grad.double().norm()) - ref_norm[n]) <= 1e-4 * ref_norm[n] + 1e-7, n
    if qtype == 'ema':
        assert C.rel_err(out['new_ema_count'], g['new_ema_count']) < 1e-6
        assert C.rel_err(out['new_ema_weight'], g['new_ema_weight']) < 1e-6
        assert C.rel_err(out['new_codebook'], g['new_codebook']) < 1e-5
    if qtype in ('standard', 'entropy'):
        assert C.rel_err(sd['quantizer.codebook.weight'].grad, g['grad_codebook']) < 1e-4


def test_oracle_matches_reference_fixture_at_256():
    """The benchmark's architecture and image size (ema_vqvae.yaml: 256 x 256, channels 128, K = 1024; batch 2) against the
    fixture made by executing the reference's modules (oracle/make_golden_256.py)."""
    g = C.golden('cfg2_256_ema')
    sd, x = C.seeded_inputs_256()
    ref_init = dict(zip(g['init_names'].tolist(), g['init_abs_sums'].tolist()))
    assert set(ref_init) == set(sd)
    for n, t in sd.items():
        assert abs(float(t.double().abs().sum()) - ref_init[n]) <= 1e-9 * max(1.0, ref_init[n]), n
    sd = oinit.make_leaf(sd, 'ema')
    torch.set_num_threads(8)
    out = orc.train_step_mse(sd, x, C.oracle_cfg('cfg2_256', 'ema'))
    assert np.array_equal(out['idx'].numpy(), g['idx']), 'codebook indices must be bit-exact'
    assert C.rel_err(out['z'], g['z']) < 1e-5
    recon = out['recon'].detach()
    assert C.rel_err(torch.nn.functional.avg_pool2d(recon, 8), g['recon_pool8']) < 2e-5
    assert C.rel_err(recon.reshape(-1)[:4096], g['recon_head']) < 2e-5
    assert abs(float(recon.double().abs().sum()) - float(g['recon_abs_sum'])) <= 1e-6 * float(g['recon_abs_sum'])
    assert abs(out['q_loss'].item() - float(g['q_loss'])) <= 2e-6
    assert abs(out['l2_loss'].item() - float(g['l2'])) <= 1e-6
    assert C.rel_err(sd['encoder.conv_in.weight'].grad, g['grad_enc_conv_in']) < 1e-4
    assert C.rel_err(sd['decoder.conv_out.weight'].grad, g['grad_dec_conv_out']) < 1e-4
    ref_norm = dict(zip(g['grad_names'].tolist(), g['grad_norms'].tolist()))
    for n, t in sd.items():
        if t.grad is not None and n in ref_norm:
            assert abs(float(t.grad.double().norm()) - ref_norm[n]) <= 1e-4 * ref_norm[n] + 1e-7, n
    assert C.rel_err(out['new_ema_count'], g['new_ema_count']) < 1e-6
    assert C.rel_err(out['new_ema_weight'].double().sum(1), g['new_ema_weight_rowsum']) < 1e-6
    assert C.rel_err(out['new_codebook'].double().sum(1), g['new_codebook_rowsum']) < 1e-5


def test_survey_known_answers():
    """The SURVEY.md 8(c) known-answer values, which were produced independently of make_golden.py."""
    g = C.golden('cfg1_standard')
    assert abs(float(g['q_loss']) - 0.16531526) < 1e-7
    assert abs(float(g['l2']) - 0.41883290) < 1e-7
    assert g['idx'][0, :8].tolist() == [124, 124, 124, 155, 3, 210, 210, 55]
    assert int(g['idx'].sum()) == 15362
    g = C.golden('cfg1_ema')
    assert g['idx'][0, :8].tolist() == [114, 107, 155, 107, 54, 87, 87, 86]
    assert int(g['idx'].sum()) == 20582
    assert C.golden('cfg1_gumbel')['idx'].shape == (8, 4, 4)          # defect B5: (B,H,W)


def test_vq_ema_kernel_shape_fixture():
    for init in ('uniform', 'normal'):
        g = C.golden(f'vqema_N4096_K1024_{init}')
        torch.manual_seed(77)
        K, D, N = 1024, 256, 4096
        cb = torch.nn.Embedding(K, D).weight.detach().clone()
        ema_w = torch.empty(K, D).uniform_(-1 / K, 1 / K)
        cb.uniform_(-1 / K, 1 / K)
        if init == 'normal':
            cb.normal_(); ema_w.copy_(cb)
        z = torch.randn(N // 256, D, 16, 16)
        q, idx, loss, ncb, ncnt, nw = orc.vq_ema(z, cb, torch.zeros(K), ema_w, 0.25, 0.95, 1e-5, True)
        assert np.array_equal(idx.reshape(-1).numpy().astype(np.int32), g['idx'])
        assert abs(loss.item() - float(g['loss'])) < 1e-6
        assert C.rel_err(ncnt, g['new_ema_count']) < 1e-6
        assert C.rel_err(ncb.double().sum(1), g['codebook_rowsum']) < 1e-5


def test_schedules_and_usage():
    assert orc.cosine_schedule(0, 0, 100, 1.0, 0.5) == 1.0
    assert abs(orc.cosine_schedule(50, 0, 100, 1.0, 0.5) - 0.75) < 1e-12
    assert orc.cosine_schedule(200, 0, 100, 1.0, 0.5) == 0.5
    assert abs(orc.linear_schedule(5, 0, 10, 0.0, 1.0) - 0.5) < 1e-12
    assert abs(orc.linear_cosine_schedule(10, 0, 110, 1.0, 0.5, 10) - 1.0) < 1e-12
    p, perp, used = orc.codebook_usage(torch.tensor([1.0, 1.0, 0.0, 2.0]))
    assert used == 75.0 and abs(perp - float(np.exp(1.5 * np.log(2)))) < 1e-5


def test_adamw_groups_collision():
    """Defect B2: encoder tensors whose relative name also exists in the decoder never reach the optimizer."""
    sd = oinit.init_state('ema', 256, 256, 128, 2, (1, 2, 2, 4), seed=0)
    enc = [n[len('encoder.'):] for n in sd if n.startswith('encoder.')]
    dec = [n[len('decoder.'):] for n in sd if n.startswith('decoder.')]
    q = ['codebook.weight']
    decay, no_decay = orc.adamw_groups(enc, dec, q, replicate_name_collision=True)
    got = set(decay) | set(no_decay)
    dropped = [n for n in sd if n.startswith('encoder.') and n not in got]
    assert len(dropped) == 53                                             # SURVEY.md 3.5 B2 [probe]
    assert sum(sd[n].numel() for n in dropped) == 13415296
    decay2, no_decay2 = orc.adamw_groups(enc, dec, q, replicate_name_collision=False)
    assert len(set(decay2) | set(no_decay2)) == len(enc) + len(dec) + 1
    assert 'quantizer.codebook.weight' in no_decay2 and 'encoder.conv_in.weight' in decay2
    assert 'encoder.norm.weight' in no_decay2 and 'decoder.blocks.2.conv.bias' in no_decay2


@pytest.mark.parametrize('loss_type', ['softmax', 'argmax'])
def test_oracle_entropy_quantizer_matches_reference_module(loss_type):
    """oracle.vq_entropy (both target types of vector_quantizers.py:296-328) against the reference module's forward+backward."""
    import numpy as np
    import torch
    from oracle import vqvae_oracle as O
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', f'quantizer_entropy_{loss_type}.npz'))
    z = torch.from_numpy(g['z']).requires_grad_(); cb = torch.from_numpy(g['codebook']).requires_grad_()
    q, idx, loss = O.vq_entropy(z, cb, 0.25, 0.1, 0.01, loss_type)
    (loss * 1.5 + (q * torch.from_numpy(g['g_q'])).sum()).backward()
    assert np.array_equal(idx.numpy(), g['idx'])
    assert abs(float(loss) - float(g['loss'])) < 1e-6
    assert np.allclose(z.grad.numpy(), g['dz'], atol=2e-5 * np.abs(g['dz']).max())
    assert np.allclose(cb.grad.numpy(), g['dcb'], atol=2e-5 * np.abs(g['dcb']).max())


def test_oracle_ssim_properties():
    """oracle.ssim_torchmetrics (restated published algorithm; torchmetrics is absent, parity unpinned): identity, symmetry,
    invariance under a common scaling (data_range=None follows the images), and the reflect-pad + crop form equals VALID
    Gaussian windows of the un-padded image -- the form the CUDA kernel evaluates."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    x = torch.rand(2, 3, 40, 36, dtype=torch.float64)
    y = (x + 0.1 * torch.randn_like(x)).clamp(0, 1)
    s = orc.ssim_torchmetrics(x, y)
    assert s.shape == (2,) and bool((s < 1).all()) and bool((s > 0).all())
    assert float((orc.ssim_torchmetrics(x, x) - 1).abs().max()) < 1e-12
    assert float((orc.ssim_torchmetrics(y, x) - s).abs().max()) < 1e-12
    assert float((orc.ssim_torchmetrics(3 * x, 3 * y) - s).abs().max()) < 1e-10
    dist = torch.arange(-5, 6, dtype=torch.float64)
    g = torch.exp(-((dist / 1.5) ** 2) / 2); g = g / g.sum()
    k = torch.outer(g, g).expand(3, 1, 11, 11)
    R = torch.maximum(x.max() - x.min(), y.max() - y.min())
    c1, c2 = (0.01 * R) ** 2, (0.03 * R) ** 2
    f = lambda t: F.conv2d(t, k, groups=3)
    mx, my = f(x), f(y)
    sx, sy, sxy = (f(x * x) - mx * mx).clamp(min=0), (f(y * y) - my * my).clamp(min=0), f(x * y) - mx * my
    direct = (((2 * mx * my + c1) * (2 * sxy + c2)) / ((mx * mx + my * my + c1) * (sx + sy + c2))).reshape(2, -1).mean(-1)
    assert float((direct - s).abs().max()) < 1e-12
