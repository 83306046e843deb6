"""GPU: the launch-folded forms must compute what the separate launches compute.

* vqb_gn_apply_sums      == vqb_gn_finalize + vqb_gn_apply            (bit-identical: same double arithmetic per group)
* vqb_gn_bwd_apply_part  == vqb_gn_bwd_finalize + vqb_gn_bwd_apply    (bit-identical dx / coef; dgamma, dbeta overwritten or +=)
* vqb_unpack_conv_wgrad_acc (accumulate / rezero) vs vqb_unpack_conv_wgrad
* a Trainer step with gradients accumulated DIRECTLY into the flat gradient buffers (ops.grad_sink) and scratch from the
  per-step zero arena == the same step with autograd's AccumulateGrad and torch.zeros (reference: model.py:244-275)."""
import pytest
import torch

from oracle import init_state as oinit
from oracle.step_cases import STEP_CASES, q_conf_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def V():
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    pkg.lib.load()
    yield pkg
    pkg.set_precision('strict')


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('n,c,h,w', [(3, 128, 16, 16), (2, 256, 8, 8), (2, 64, 5, 7), (1, 512, 4, 4)])
def test_gn_folded_equals_separate(V, dtype, n, c, h, w):
    from vqvae_vqgan_pytorch_lightning_b200.lib import ACT_SILU, call, dt, ptr, stream
    torch.manual_seed(3)
    G, eps = 32, 1e-6
    x = (torch.randn(n, c, h, w, device='cuda') * 2 + 0.5).to(dtype).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(n, c, h, w, device='cuda').to(dtype).contiguous(memory_format=torch.channels_last)
    skip = torch.randn(n, c, h, w, device='cuda').to(dtype).contiguous(memory_format=torch.channels_last)
    ga, be = torch.rand(c, device='cuda') + 0.5, torch.randn(c, device='cuda') * 0.1
    sums = torch.zeros(n * G * 2, dtype=torch.float64, device='cuda')
    call('vqb_gn_stats', ptr(x), dt(x), ptr(sums), n, h * w, c, G, stream())
    # separate launches
    stats = torch.empty(n * G * 2, device='cuda')
    call('vqb_gn_finalize', ptr(sums), ptr(stats), n, h * w, c, G, eps, stream())
    y = torch.empty_like(x)
    call('vqb_gn_apply', ptr(x), dt(x), ptr(stats), ptr(ga), ptr(be), ptr(y), dt(y), n, h * w, c, G, ACT_SILU, stream())
    # folded
    stats2 = torch.full((n * G * 2,), float('nan'), device='cuda')
    y2 = torch.empty_like(x)
    call('vqb_gn_apply_sums', ptr(x), dt(x), ptr(sums), ptr(ga), ptr(be), ptr(y2), dt(y2), ptr(stats2), n, h * w, c, G, eps, ACT_SILU, stream())
    assert torch.equal(stats, stats2)
    assert torch.equal(y, y2)
    # backward
    part = torch.zeros(n * c * 2, dtype=torch.float64, device='cuda')
    call('vqb_gn_bwd_reduce', ptr(x), dt(x), ptr(dy), dt(dy), ptr(stats), ptr(ga), ptr(be), ptr(part), n, h * w, c, G, ACT_SILU, stream())
    coef = torch.empty(n * G * 2, device='cuda'); dga = torch.empty(c, device='cuda'); dbe = torch.empty(c, device='cuda')
    call('vqb_gn_bwd_finalize', ptr(part), ptr(ga), ptr(coef), ptr(dga), ptr(dbe), n, h * w, c, G, stream())
    for add in (None, skip):
        dx = torch.empty_like(x)
        call('vqb_gn_bwd_apply', ptr(x), dt(x), ptr(dy), dt(dy), ptr(stats), ptr(ga), ptr(be), ptr(coef), ptr(add), ptr(dx), dt(dx),
             n, h * w, c, G, ACT_SILU, stream())
        for acc in (0, 1):
            dx2 = torch.empty_like(x)
            dga2 = torch.full((c,), 2.0, device='cuda'); dbe2 = torch.full((c,), -1.0, device='cuda')
            call('vqb_gn_bwd_apply_part', ptr(x), dt(x), ptr(dy), dt(dy), ptr(stats), ptr(ga), ptr(be), ptr(part), ptr(add), ptr(dx2), dt(dx2),
                 ptr(dga2), ptr(dbe2), acc, n, h * w, c, G, ACT_SILU, stream())
            assert torch.equal(dx, dx2)
            if acc:
                assert torch.allclose(dga2, dga + 2.0, rtol=1e-6, atol=1e-6) and torch.allclose(dbe2, dbe - 1.0, rtol=1e-6, atol=1e-6)
            else:
                assert torch.equal(dga2, dga) and torch.equal(dbe2, dbe)


def test_unpack_wgrad_accumulate_rezero(V):
    from vqvae_vqgan_pytorch_lightning_b200.lib import call, ptr, stream
    torch.manual_seed(4)
    co, ci, k = 70, 45, 3
    dwp = torch.randn(k * k * ci * co, device='cuda')
    ref = torch.empty(co, ci, k, k, device='cuda')
    call('vqb_unpack_conv_wgrad', ptr(dwp), ptr(ref), co, ci, k, k, 0.5, stream())
    base = torch.randn(co, ci, k, k, device='cuda')
    for acc in (0, 1):
        for rz in (0, 1):
            src, dst = dwp.clone(), base.clone()
            call('vqb_unpack_conv_wgrad_acc', ptr(src), ptr(dst), co, ci, k, k, 0.5, acc, rz, stream())
            assert torch.equal(dst, base + ref if acc else ref)
            assert torch.equal(src, torch.zeros_like(src) if rz else dwp)


def _grads_of_one_step(V, name, mode, direct):
    """flat gradient buffers of every optimizer at the moment its step() is called in ONE training step (the update itself is
    skipped, so both runs differentiate the same weights)"""
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    V.set_precision(mode)
    case = dict(STEP_CASES[name])
    if name.startswith('gan'):
        case['l_conf'] = dict(case['l_conf'], adversarial_params=dict(case['l_conf']['adversarial_params'], r1_reg_every=1))
    crit = None if case['l_conf'] is None else ('gan' if case['l_conf']['adversarial_params'] is not None else 'lpips')
    sd = oinit.init_state(case['qtype'], case['K'], case['D'], case['ch'], case['nrb'], case['mult'], seed=case['seed'],
                          criterion=crit, image_size=case['S'])
    model = V.VQVAE(case['S'], dict(channels=case['ch'], num_res_blocks=case['nrb'], channel_multipliers=list(case['mult'])),
                    q_conf_of(case), case['l_conf'], dict(case['t_conf']), pretrained_lpips=False)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    model.training_augmentations = None
    tr = Trainer(max_epochs=1, num_training_batches=4)
    tr.attach(model)
    flagged = [p for p in model.parameters() if hasattr(p, '_vqb_direct_grad')]
    assert flagged and all(p._vqb_direct_grad for p in flagged)       # single process: Trainer.attach opts every optimizer tensor in
    for p in flagged:
        p._vqb_direct_grad = direct
    grads = []
    for o in tr.optimizers:
        o.step = (lambda o=o: grads.append(o.flat_grad.clone()))          # record, do not update
    model.on_train_start()
    torch.manual_seed(case['seed'] + 1)
    x = torch.rand(case['B'], 3, case['S'], case['S']).cuda()
    arena = V.ops.zero_arena
    if not direct:
        V.ops.zero_arena = V.ops.ZeroArena(0)              # a pool of 0 bytes: every request falls through to torch.zeros
    try:
        loss = float(tr.run_step(x, 0))
    finally:
        V.ops.zero_arena = arena
    torch.cuda.synchronize()
    return grads, loss


@pytest.mark.parametrize('name,mode', [('mse_ema', 'strict'), ('mse_ema', 'fast'), ('mse_gumbel', 'strict'), ('lpips_ema', 'fast'),
                                       ('gan_hinge_adaptive_r1', 'strict'), ('gan_nonsat_fixed', 'fast')])
def test_direct_gradient_accumulation_equals_autograd(V, name, mode):
    if mode == 'fast' and not V.lib.load().vqb_device_supports_tcgen05():
        pytest.skip('needs sm_100')
    if name == 'mse_gumbel':
        torch.manual_seed(11); torch.cuda.manual_seed(11)
    ga, la = _grads_of_one_step(V, name, mode, True)
    if name == 'mse_gumbel':
        torch.manual_seed(11); torch.cuda.manual_seed(11)
    gb, lb = _grads_of_one_step(V, name, mode, False)
    assert len(ga) == len(gb) and len(ga) >= 1
    assert abs(la - lb) <= 1e-5 * max(abs(lb), 1.0)
    # same kernels, same weights: the gradients differ by the order of floating-point atomics only (split-K weight gradients)
    for a, b in zip(ga, gb):
        assert float(b.abs().max()) > 0
        d = float((a.double() - b.double()).norm() / b.double().norm())
        print(f'{name}/{mode}: |g| {float(b.norm()):.3e} relative difference {d:.2e}')
        assert d < 1e-4, d


@pytest.mark.parametrize('n,c,h,w,with_add', [(3, 128, 128, 192, True), (5, 128, 256, 256, False), (2, 256, 128, 128, True)])
def test_gn_backward_cooperative_equals_two_launches(V, n, c, h, w, with_add):
    """vqb_gn_bwd_fused (one cooperative launch, image b-1 applied out of L2 while image b is reduced) == vqb_gn_bwd_reduce +
    vqb_gn_bwd_apply_part: same per-element arithmetic; the double-precision sums are combined in a different order."""
    from vqvae_vqgan_pytorch_lightning_b200.lib import ACT_SILU, BF16, call, dt, ptr, stream
    L = V.lib.load()
    G, eps = 32, 1e-6
    torch.manual_seed(9)
    mk = lambda: torch.randn(n, c, h, w, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last)
    x, dy, skip = mk() * 2 + 0.5, mk(), mk()
    ga, be = torch.rand(c, device='cuda') + 0.5, torch.randn(c, device='cuda') * 0.1
    sums = torch.zeros(n * G * 2, dtype=torch.float64, device='cuda')
    call('vqb_gn_stats', ptr(x), dt(x), ptr(sums), n, h * w, c, G, stream())
    stats = torch.empty(n * G * 2, device='cuda')
    call('vqb_gn_finalize', ptr(sums), ptr(stats), n, h * w, c, G, eps, stream())
    add = skip if with_add else None
    part = torch.zeros(n * c * 2, dtype=torch.float64, device='cuda')
    call('vqb_gn_bwd_reduce', ptr(x), dt(x), ptr(dy), dt(dy), ptr(stats), ptr(ga), ptr(be), ptr(part), n, h * w, c, G, ACT_SILU, stream())
    dx = torch.empty_like(x); dga = torch.empty(c, device='cuda'); dbe = torch.empty(c, device='cuda')
    call('vqb_gn_bwd_apply_part', ptr(x), dt(x), ptr(dy), dt(dy), ptr(stats), ptr(ga), ptr(be), ptr(part), ptr(add), ptr(dx), dt(dx),
         ptr(dga), ptr(dbe), 0, n, h * w, c, G, ACT_SILU, stream())
    for max_ctas in (0, 64):
        for acc in (0, 1):
            part2 = torch.zeros(n * c * 2, dtype=torch.float64, device='cuda')
            counters = torch.zeros(n, dtype=torch.int32, device='cuda')
            dx2 = torch.full_like(x, float('nan'))
            dga2 = torch.full((c,), 1.5, device='cuda'); dbe2 = torch.full((c,), -0.5, device='cuda')
            call('vqb_gn_bwd_fused', ptr(x), ptr(dy), ptr(stats), ptr(ga), ptr(be), ptr(part2), ptr(counters), ptr(add), ptr(dx2),
                 ptr(dga2), ptr(dbe2), acc, n, h * w, c, G, ACT_SILU, max_ctas, stream())
            torch.cuda.synchronize()
            assert torch.isfinite(dx2.float()).all()
            # per-thread partial sums are fp32 over DIFFERENT pixel subsets in the two forms (only the cross-CTA combine is double)
            assert float((part2 - part).norm() / part.norm()) < 1e-5
            assert float((dx2.float() - dx.float()).abs().max()) <= 2e-2 * float(dx.float().abs().max())       # a bf16 ulp where a coefficient moved by one fp32 ulp
            assert float((dx2.float() - dx.float()).norm() / dx.float().norm()) < 1e-3
            ref_g, ref_b = (dga + 1.5, dbe - 0.5) if acc else (dga, dbe)
            assert torch.allclose(dga2, ref_g, rtol=1e-4, atol=1e-3) and torch.allclose(dbe2, ref_b, rtol=1e-4, atol=1e-3)
