"""GPU parity tests: every libvqgan_b200 kernel (called through the C ABI via ctypes) against the CPU oracle
(oracle/vqvae_oracle.py, plain fp32 torch) on identical seeded inputs.  Strict (fp32 SIMT) path: tolerance
1e-4 relative (BASELINE.json north_star); indices bit-exact up to the oracle's own fp32 ties."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vqvae_oracle as orc
from tests import common as C

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope='module')
def V():
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    pkg.lib.load()
    pkg.set_precision('strict')
    return pkg


def cl(t):
    return t.cuda().contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize('n,h,w,ci,co,k,bias,res,act', [
    (2, 16, 16, 3, 32, 3, False, False, 0),
    (2, 16, 16, 32, 64, 3, False, True, 0),
    (3, 9, 7, 32, 32, 3, True, False, 0),          # ragged spatial size, M not a multiple of the tile
    (2, 8, 8, 64, 32, 1, True, False, 0),
    (2, 16, 16, 32, 3, 3, True, False, 1),         # tanh epilogue, Co=3 (decoder.conv_out)
    (1, 4, 4, 128, 256, 3, False, True, 0),
    (2, 32, 32, 128, 128, 3, False, False, 0),
])
def test_conv2d_fwd_bwd(V, n, h, w, ci, co, k, bias, res, act):
    torch.manual_seed(0)
    x = torch.randn(n, ci, h, w)
    wt = torch.randn(co, ci, k, k) / (ci * k * k) ** 0.5
    b = torch.randn(co) if bias else None
    r = torch.randn(n, co, h, w) if res else None
    go = torch.randn(n, co, h, w)
    # oracle
    xo, wo = x.clone().requires_grad_(), wt.clone().requires_grad_()
    bo = b.clone().requires_grad_() if bias else None
    ro = r.clone().requires_grad_() if res else None
    y = F.conv2d(xo, wo, bo, padding=k // 2)
    if act == 1:
        y = torch.tanh(y)
    if res:
        y = y + ro
    y.backward(go)
    # ours
    xg, wg = cl(x).requires_grad_(), wt.cuda().requires_grad_()
    bg = b.cuda().requires_grad_() if bias else None
    rg = cl(r).requires_grad_() if res else None
    yg = V.ops.conv2d(xg, wg, bg, rg, pad=k // 2, act=act)
    yg.backward(cl(go))
    assert C.rel_err(yg, y) < TOL
    assert C.rel_err(xg.grad, xo.grad) < TOL
    assert C.rel_err(wg.grad, wo.grad) < TOL
    if bias:
        assert C.rel_err(bg.grad, bo.grad) < TOL
    if res:
        assert C.rel_err(rg.grad, ro.grad) < TOL


@pytest.mark.parametrize('n,c,h,w', [(2, 32, 8, 8), (2, 64, 5, 7), (3, 128, 16, 16), (2, 256, 8, 8), (1, 512, 4, 4)])
def test_groupnorm_silu_fwd_bwd(V, n, c, h, w):
    torch.manual_seed(1)
    x = torch.randn(n, c, h, w) * 2 + 0.5
    ga, be = torch.rand(1, c, 1, 1) + 0.5, torch.randn(1, c, 1, 1) * 0.1
    go = torch.randn(n, c, h, w)
    xo, gao, beo = x.clone().requires_grad_(), ga.clone().requires_grad_(), be.clone().requires_grad_()
    y = F.silu(orc.group_norm(xo, gao, beo))
    y.backward(go)
    xg, gag, beg = cl(x).requires_grad_(), ga.cuda().requires_grad_(), be.cuda().requires_grad_()
    yg = V.ops.group_norm_act(xg, gag, beg, 32, 1e-6, V.lib.ACT_SILU)
    yg.backward(cl(go))
    assert C.rel_err(yg, y) < TOL
    assert C.rel_err(xg.grad, xo.grad) < TOL
    assert C.rel_err(gag.grad, gao.grad) < TOL
    assert C.rel_err(beg.grad, beo.grad) < TOL


def test_resample(V):
    torch.manual_seed(2)
    x = torch.randn(2, 32, 8, 8)
    go_d, go_u = torch.randn(2, 32, 4, 4), torch.randn(2, 32, 16, 16)
    xo = x.clone().requires_grad_()
    yd = F.avg_pool2d(xo, 2, 2, 0); yd.backward(go_d); gd = xo.grad.clone(); xo.grad = None
    yu = F.interpolate(xo, scale_factor=2.0, mode='nearest-exact'); yu.backward(go_u); gu = xo.grad.clone()
    xg = cl(x).requires_grad_()
    ydg = V.ops.avg_pool2(xg); ydg.backward(cl(go_d)); gdg = xg.grad.clone(); xg.grad = None
    yug = V.ops.upsample2(xg); yug.backward(cl(go_u)); gug = xg.grad.clone()
    for a, b in ((ydg, yd), (gdg, gd), (yug, yu), (gug, gu)):
        assert C.rel_err(a, b) < 1e-6


def test_layout_and_losses(V):
    torch.manual_seed(3)
    img = torch.rand(2, 3, 8, 8) * 1.4 - 0.2            # values outside [0,1] exercise the clamp
    ref = orc.normalize_images(img)
    ours = V.ops.images_to_nhwc(img.cuda(), torch.float32)
    assert ours.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(ours.cpu(), ref)
    back = V.ops.nhwc_to_images(ours, 0.5, 0.5, (0.0, 1.0))
    assert C.rel_err(back, torch.clamp(img, 0, 1)) < 1e-6
    wide = torch.randn(2, 40, 5, 6)
    assert torch.equal(V.ops.images_to_nhwc(wide.cuda(), torch.float32, normalize=False).cpu(), wide)
    assert torch.equal(V.ops.nhwc_to_images(cl(wide)).cpu(), wide)
    a, b = torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8)
    ao = a.clone().requires_grad_()
    lo = 0.7 * F.mse_loss(ao, b) + 0.3 * (ao - b).abs().mean()
    lo.backward()
    ag = cl(a).requires_grad_()
    l2, l1 = V.ops.mse_l1(ag, cl(b))
    (0.7 * l2 + 0.3 * l1).backward()
    assert abs(float(l2) - float(F.mse_loss(a, b))) < 1e-6
    assert C.rel_err(ag.grad, ao.grad) < 1e-5


@pytest.mark.parametrize('init', ['uniform', 'normal'])
def test_vq_assign_matches_reference_fixture(V, init):
    """N=4096, K=1024, D=256 against the fixture produced by the reference's EMAVectorQuantizer."""
    g = C.golden(f'vqema_N4096_K1024_{init}')
    torch.manual_seed(77)
    K, D, N = 1024, 256, 4096
    cb = torch.nn.Embedding(K, D).weight.detach().clone()
    ema_w = torch.empty(K, D).uniform_(-1 / K, 1 / K)
    cb.uniform_(-1 / K, 1 / K)
    if init == 'normal':
        cb.normal_(); ema_w.copy_(cb)
    z = torch.randn(N // 256, D, 16, 16)
    from vqvae_vqgan_pytorch_lightning_b200.modules.vector_quantizers import EMAVectorQuantizer
    q = EMAVectorQuantizer(K, D, 0.25, 0.95, 1e-5).cuda()
    with torch.no_grad():
        q.codebook.weight.copy_(cb); q.ema_weight.copy_(ema_w)
    q.train()
    quant, idx, loss = q(cl(z))
    flat = z.permute(0, 2, 3, 1).reshape(N, D)
    exact, ties, bad = C.tie_aware_index_check(idx, g['idx'], flat, cb)
    assert bad == 0, (exact, ties, bad)
    if init == 'normal':
        assert ties == 0                                      # tie-free codebook: bit-exact indices
    assert abs(float(loss) - float(g['loss'])) < 1e-5 * max(1.0, abs(float(g['loss'])))
    if ties == 0:
        assert C.rel_err(q.ema_count, g['new_ema_count']) < 1e-5
        assert C.rel_err(q.codebook.weight.double().sum(1), g['codebook_rowsum']) < 1e-4
        assert abs(float(quant.double().sum()) - float(g['quant_sum'])) < 1e-3 * max(1.0, abs(float(g['quant_sum'])))


def test_vq_edge_cases(V):
    """ragged N (not a multiple of the 64-row tile), K not a multiple of 128, duplicate codes (first index wins)."""
    torch.manual_seed(5)
    N, K, D = 100, 37, 32
    z = torch.randn(N, D)
    cb = torch.randn(K, D)
    cb[20] = cb[5]                                             # exact duplicate: argmin must return 5, never 20
    q, idx, sse, counts, dw = V.ops.vq_assign_raw(z.cuda(), cb.cuda(), 0, True, True)
    d = orc.l2_distances(z, cb)
    ref = torch.argmin(d, dim=1)
    exact, ties, bad = C.tie_aware_index_check(idx, ref, z, cb)
    assert bad == 0
    assert not (idx == 20).any()
    assert float(counts.sum()) == N
    ref_dw = torch.zeros(K, D).index_add_(0, idx.cpu(), z)
    assert C.rel_err(dw, ref_dw) < 1e-5
    assert C.rel_err(q, z + (cb[idx.cpu()] - z)) < 1e-6
    with pytest.raises(V.lib.VQBError):
        V.ops.vq_assign_raw(torch.randn(4, 30).cuda(), torch.randn(8, 30).cuda(), 0)     # D % 4 != 0 is rejected loudly


def test_adamw(V):
    torch.manual_seed(6)
    n = 10007
    p, g = torch.randn(n), torch.randn(n)
    m, v = torch.zeros(n), torch.zeros(n)
    pg, gg, mg, vg = p.cuda(), g.cuda(), m.cuda(), v.cuda()
    for step in (1, 2, 3):
        orc.adamw_step(p, g, m, v, step, 1e-3, 0.0, 0.99, 1e-8, 1e-4)
        V.ops.adamw_flat(pg, gg, mg, vg, 1e-3, 0.0, 0.99, 1e-8, 1e-4, step)
    assert C.rel_err(pg, p) < 1e-6 and C.rel_err(vg, v) < 1e-6


def _build_model(V, case, qtype, sd):
    c = C.CASES[case]
    qp = {k: v for k, v in C.Q_PARAMS[qtype].items() if k != 'type'}
    if qtype == 'gumbel':
        qp.update(kl_warmup_epochs=None, temp_decay_epochs=None, temp_final=None)
    model = V.VQVAE(c['S'], dict(channels=c['ch'], num_res_blocks=c['nrb'], channel_multipliers=list(c['mult'])),
                    dict(num_embeddings=c['K'], embedding_dim=c['D'], type=qtype, params=qp, reinit_every_n_epochs=None),
                    None, dict(lr=1e-4, betas=[0.0, 0.99], eps=1e-8, weight_decay=1e-4, warmup_epochs=None, decay_epochs=None))
    model.load_state_dict(sd)
    model.training_augmentations = None                    # the fixtures were produced without augmentation
    return model.cuda().train()


@pytest.mark.parametrize('case', ['tiny', 'cfg1'])
@pytest.mark.parametrize('qtype', ['standard', 'ema'])
def test_train_step_matches_reference_fixture(V, case, qtype):
    """Whole forward + backward of the MSE branch against fixtures produced by the reference's own modules."""
    g = C.golden(f'{case}_{qtype}')
    sd, x = C.seeded_inputs(case, qtype)
    model = _build_model(V, case, qtype, sd)
    xg = cl(x)
    recon, q_loss, idx = model(xg)
    l2 = model.criterion(recon, xg)
    (q_loss + l2).backward()
    c = C.CASES[case]
    z = model.encoder(xg).detach()
    assert C.rel_err(z, g['z']) < TOL
    flat = torch.from_numpy(g['z']).permute(0, 2, 3, 1).reshape(-1, c['D'])
    exact, ties, bad = C.tie_aware_index_check(idx, g['idx'], flat, sd['quantizer.codebook.weight'])
    assert bad == 0 and ties <= max(1, idx.numel() // 100), (exact, ties, bad)
    if ties == 0:
        assert C.rel_err(recon, g['recon']) < TOL
        assert abs(float(q_loss) - float(g['q_loss'])) <= TOL * max(1.0, abs(float(g['q_loss'])))
        assert abs(float(l2) - float(g['l2'])) <= TOL
        assert C.rel_err(model.encoder.conv_in.weight.grad, g['grad_enc_conv_in']) < 5 * TOL
        assert C.rel_err(model.decoder.conv_out.weight.grad, g['grad_dec_conv_out']) < TOL
        assert C.rel_err(model.decoder.conv_out.bias.grad, g['grad_dec_conv_out_bias']) < TOL
        ref_norm = dict(zip(g['grad_names'].tolist(), g['grad_norms'].tolist()))
        for n, p in model.named_parameters():
            if p.grad is not None and n in ref_norm:
                assert abs(float(p.grad.double().norm()) - ref_norm[n]) <= 5 * TOL * ref_norm[n] + 1e-7, n
        if qtype == 'ema':
            assert C.rel_err(model.quantizer.ema_count, g['new_ema_count']) < 2e-5
            assert C.rel_err(model.quantizer.ema_weight, g['new_ema_weight']) < 2e-5        # sums of z rows (z itself: 1e-4 bar)
        else:
            assert C.rel_err(model.quantizer.codebook.weight.grad, g['grad_codebook']) < TOL


@pytest.mark.parametrize('case', ['tiny', 'cfg1'])
def test_entropy_train_step_matches_reference_fixture(V, case):
    """Entropy quantizer (cfg3 family): forward + backward incl. the softmax-entropy regulariser and its codebook gradient."""
    g = C.golden(f'{case}_entropy')
    sd, x = C.seeded_inputs(case, 'entropy')
    model = _build_model(V, case, 'entropy', sd)
    xg = cl(x)
    recon, q_loss, idx = model(xg)
    l2 = model.criterion(recon, xg)
    (q_loss + l2).backward()
    c = C.CASES[case]
    flat = torch.from_numpy(g['z']).permute(0, 2, 3, 1).reshape(-1, c['D'])
    exact, ties, bad = C.tie_aware_index_check(idx, g['idx'], flat, sd['quantizer.codebook.weight'], order='entropy')
    assert bad == 0 and ties <= max(1, idx.numel() // 100), (exact, ties, bad)
    if ties == 0:
        assert abs(float(q_loss) - float(g['q_loss'])) <= TOL * max(1.0, abs(float(g['q_loss'])))
        assert C.rel_err(recon, g['recon']) < TOL
        assert C.rel_err(model.quantizer.codebook.weight.grad, g['grad_codebook']) < 5 * TOL
        assert C.rel_err(model.encoder.conv_in.weight.grad, g['grad_enc_conv_in']) < 5 * TOL
        ref_norm = dict(zip(g['grad_names'].tolist(), g['grad_norms'].tolist()))
        for n, p in model.named_parameters():
            if p.grad is not None and n in ref_norm:
                assert abs(float(p.grad.double().norm()) - ref_norm[n]) <= 5 * TOL * ref_norm[n] + 1e-7, n


@pytest.mark.parametrize('case', ['tiny', 'cfg1'])
def test_gumbel_train_step_matches_reference_fixture(V, case):
    """Gumbel quantizer (cfg4 family) with the reference's own Exp(1) noise fed explicitly to the kernel."""
    g = C.golden(f'{case}_gumbel')
    sd, x = C.seeded_inputs(case, 'gumbel')
    model = _build_model(V, case, 'gumbel', sd)
    xg = cl(x)
    noise = cl(torch.from_numpy(g['exp_noise']))
    z = model.encoder(xg)
    quant, idx, q_loss = model.quantizer(z, exp_noise=noise)
    recon = model.decoder(quant)
    l2 = model.criterion(recon, xg)
    (q_loss + l2).backward()
    assert tuple(idx.shape) == tuple(g['idx'].shape)                     # (B,H,W): defect B5 replicated
    assert C.rel_err(z, g['z']) < TOL
    mism = int((idx.cpu() != torch.from_numpy(g['idx'])).sum())
    assert mism <= max(1, idx.numel() // 100)
    assert abs(float(q_loss) - float(g['q_loss'])) <= TOL * max(1e-3, abs(float(g['q_loss'])))
    assert C.rel_err(quant, g['quantized']) < TOL
    assert C.rel_err(recon, g['recon']) < TOL
    assert abs(float(l2) - float(g['l2'])) <= TOL
    assert C.rel_err(model.quantizer.x_to_logits.weight.grad, g['grad_x_to_logits']) < 5 * TOL
    assert C.rel_err(model.encoder.conv_in.weight.grad, g['grad_enc_conv_in']) < 5 * TOL
    ref_norm = dict(zip(g['grad_names'].tolist(), g['grad_norms'].tolist()))
    for n, p in model.named_parameters():
        if p.grad is not None and n in ref_norm:
            assert abs(float(p.grad.double().norm()) - ref_norm[n]) <= 5 * TOL * ref_norm[n] + 1e-7, n


@pytest.mark.parametrize('loss_type', ['softmax', 'argmax'])
def test_entropy_quantizer_matches_reference_module_fixture(V, loss_type):
    """EntropyVectorQuantizer forward + backward, both target types (vector_quantizers.py:296-328: softmax targets, and
    straight-through one-hot 'argmax' targets), against the reference module run on the same inputs."""
    from vqvae_vqgan_pytorch_lightning_b200.modules.vector_quantizers import EntropyVectorQuantizer
    g = C.golden(f'quantizer_entropy_{loss_type}')
    K, D = g['codebook'].shape
    q = EntropyVectorQuantizer(K, D, ent_loss_ratio=0.1, ent_temperature=0.01, ent_loss_type=loss_type, commitment_cost=0.25).cuda()
    with torch.no_grad():
        q.codebook.weight.copy_(torch.from_numpy(g['codebook']))
    z = torch.from_numpy(g['z']).cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    out, idx, loss = q(z)
    (loss * 1.5 + (out * torch.from_numpy(g['g_q']).cuda()).sum()).backward()
    assert torch.equal(idx.cpu(), torch.from_numpy(g['idx']))
    assert C.rel_err(out, g['q']) < 1e-6
    assert abs(float(loss) - float(g['loss'])) < 1e-5 * max(1.0, abs(float(g['loss'])))
    assert C.rel_err(z.grad, g['dz']) < 1e-4
    assert C.rel_err(q.codebook.weight.grad, g['dcb']) < 1e-4


def test_test_step_metrics_match_definitions(V):
    """test_step / on_test_epoch_end (model.py:491-562): MSE, PSNR (torchmetrics definitions) and accumulated codebook usage
    against a direct torch evaluation of the same reconstructions."""
    import math
    torch.manual_seed(9)
    model = V.VQVAE(32, dict(channels=32, num_res_blocks=1, channel_multipliers=[1, 2]),
                    dict(num_embeddings=32, embedding_dim=32, type='standard', params=dict(commitment_cost=0.25), reinit_every_n_epochs=None),
                    None, dict(lr=1e-4, betas=[0.9, 0.99], eps=1e-8, weight_decay=0.0, warmup_epochs=None, decay_epochs=None)).cuda().eval()
    batches = [torch.rand(4, 3, 32, 32, device='cuda') * 0.8 + 0.1 for _ in range(3)]
    model.on_test_epoch_start()
    for b in batches:
        model.test_step(b, 0)
    model.on_test_epoch_end()
    rec = torch.cat([model.reconstruct(b) for b in batches]); tgt = torch.cat(batches)
    mse = float(((rec - tgt).double() ** 2).mean())
    rng = float(tgt.max().clamp(min=0) - tgt.min().clamp(max=0))          # torchmetrics' PSNR(data_range=None) starts min and max at 0
    assert abs(float(model.logged['mse']) - mse) < 1e-6 * max(mse, 1e-6) + 1e-9
    assert abs(float(model.logged['psnr']) - 10 * math.log10(rng * rng / mse)) < 1e-3
    usage = sum(torch.bincount(model.get_tokens(b).view(-1), minlength=32) for b in batches)
    assert torch.equal(model.test_usage_count, usage) and int(usage.sum()) == 12 * 64
    assert 0 < float(model.logged['perplexity']) <= 32
    # SSIM: mean over all images of the per-image values, each batch with its own data range (torchmetrics' update / compute)
    ssim = torch.cat([orc.ssim_torchmetrics(model.reconstruct(b).cpu(), b.cpu()) for b in batches]).mean()
    assert abs(float(model.logged['ssim']) - float(ssim)) < 2e-5


@pytest.mark.parametrize('shape', [(2, 3, 32, 32), (3, 3, 64, 48), (1, 1, 11, 11), (2, 3, 256, 256), (5, 2, 27, 45)])
def test_ssim_kernel_matches_published_definition(V, shape):
    """vqb_ssim_sums against the restated torchmetrics algorithm (oracle.ssim_torchmetrics: reflect pad + depth-wise Gaussian
    conv + crop), close image pairs and unrelated ones, ragged tile edges, the smallest legal image (one window)."""
    from vqvae_vqgan_pytorch_lightning_b200 import ops
    torch.manual_seed(3)
    tgt = torch.rand(*shape)
    for noise in (0.0, 0.05, 1.0):
        rec = (tgt + noise * torch.randn(*shape)).clamp(0, 1) if noise < 1.0 else torch.rand(*shape)
        got = ops.ssim_per_image(rec.cuda(), tgt.cuda()).cpu()
        ref = orc.ssim_torchmetrics(rec.double(), tgt.double())
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) < 2e-5, (shape, noise, got, ref)
        if noise == 0.0:
            assert float((got - 1.0).abs().max()) < 1e-6


def test_batched_weight_pack_equals_single_packs(V):
    """vqb_pack_conv_weights_batched (one launch for every kernel-layout weight copy of a model) against vqb_pack_conv_weight, all
    six layouts, bf16 and fp32 outputs, through the registry the convolutions use (re-pack on a weights-epoch bump)."""
    from vqvae_vqgan_pytorch_lightning_b200 import ops
    from vqvae_vqgan_pytorch_lightning_b200.lib import call, ptr, dt, stream
    torch.manual_seed(12)
    shapes = [(128, 64, 3, 3), (256, 128, 1, 1), (64, 64, 3, 3), (128, 3, 3, 3), (3, 128, 3, 3), (70, 130, 3, 3)]
    ws = [torch.nn.Parameter(torch.randn(*s).cuda()) for s in shapes]
    combos = []
    for w in ws:
        co, ci, kh, kw = w.shape
        for mode, dtype in ((0, torch.float32), (1, torch.float32), (2, torch.bfloat16), (3, torch.bfloat16)):
            combos.append((w, mode, dtype, 1.0 if mode != 2 else 0.5))
        if kh * kw * ci <= 64:
            combos.append((w, 4, torch.bfloat16, 1.0))
        if kh * kw * co <= 64:
            combos.append((w, 5, torch.bfloat16, 1.0))
    for rnd in range(3):
        got = [ops._packed_weight(w, m, d, sc) for (w, m, d, sc) in combos]      # round 0 registers, rounds 1-2 use the batched launch
        for (w, m, d, sc), g in zip(combos, got):
            co, ci, kh, kw = w.shape
            ref = torch.empty_like(g)
            call('vqb_pack_conv_weight', ptr(w.detach()), ptr(ref), dt(ref), m, co, ci, kh, kw, sc, stream())
            assert torch.equal(g, ref), (rnd, tuple(w.shape), m)
        with torch.no_grad():
            for w in ws:
                w.data.mul_(1.01)                    # parameters rewritten behind autograd's back, as the optimizer kernel does
        ops.bump_weights_epoch()
