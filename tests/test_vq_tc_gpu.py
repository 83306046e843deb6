"""GPU: the tensor-core nearest-code search (bf16 hi/lo split + exact fp32 re-evaluation of near-ties) must return the
same indices as the exact fp32 kernel -- bit-exact for tie-free codebooks, and via the exact re-rank for the reference's
tie-heavy U(+-1/K) initial codebook.  Both tensor-core forms are covered: the one-launch fused kernel (vqb_vq_fused, the
product path) and the round-1 multi-launch path (vqb_vq_assign_tc, kept for A/B measurements)."""
import pytest
import torch

from oracle import vqvae_oracle as orc
from tests import common as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def V():
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    pkg.lib.load()
    if not pkg.lib.load().vqb_device_supports_tcgen05():
        pytest.skip('needs sm_100')
    return pkg


@pytest.mark.parametrize('mode', ['fused', 'legacy'])
@pytest.mark.parametrize('N,K,D,init', [(4096, 1024, 256, 'normal'), (4096, 1024, 256, 'uniform'), (1000, 512, 64, 'normal'),
                                        (16384, 1024, 256, 'normal'), (2048, 8192, 256, 'normal'), (300, 264, 128, 'trained'),
                                        (129, 8, 192, 'normal'), (257, 40, 64, 'uniform')])
def test_tc_search_equals_exact_kernel(V, N, K, D, init, mode):
    torch.manual_seed(9)
    z = torch.randn(N, D).cuda()
    if init == 'uniform':
        cb = torch.empty(K, D).uniform_(-1 / K, 1 / K).cuda()
    elif init == 'trained':
        cb = (z[torch.randint(0, N, (K,))] + 0.05 * torch.randn(K, D, device='cuda'))      # codes sit near the data
    else:
        cb = torch.randn(K, D).cuda()
    q0, i0, s0, c0, w0 = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=False)
    q1, i1, s1, c1, w1 = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=mode)
    und = int(V.ops.vq_assign_raw.last_undecided)
    assert torch.equal(i0, i1), (int((i0 != i1).sum()), und)
    assert torch.equal(q0, q1)
    assert abs(float(s0) - float(s1)) <= 1e-6 * abs(float(s0))
    assert torch.equal(c0, c1) and C.rel_err(w1, w0) < 1e-5
    if init == 'normal':
        assert und <= (N // 20 if mode == 'legacy' else N // 8), und      # tie-free: most rows are decided by the tensor-core pass
    if init == 'uniform':
        assert und > 0                             # tie-heavy: the exact path takes over for the near-tied rows
    # and against the CPU oracle, tie-aware
    ref = torch.argmin(orc.l2_distances(z.cpu(), cb.cpu()), dim=1)
    exact, ties, bad = C.tie_aware_index_check(i1, ref, z.cpu(), cb.cpu())
    assert bad == 0


def test_tc_search_entropy_order_and_fixture(V):
    g = C.golden('vqema_N4096_K1024_normal')
    torch.manual_seed(77)
    K, D, N = 1024, 256, 4096
    cb = torch.nn.Embedding(K, D).weight.detach().clone()
    _ = torch.empty(K, D).uniform_(-1 / K, 1 / K)
    cb.uniform_(-1 / K, 1 / K); cb.normal_()
    z = torch.randn(N // 256, D, 16, 16)
    flat = z.permute(0, 2, 3, 1).reshape(N, D).contiguous().cuda()
    for mode in ('fused', 'legacy'):
        for order in (0, 1):
            _, idx, _, _, _ = V.ops.vq_assign_raw(flat, cb.cuda(), order, False, False, use_tc=mode)
            assert torch.equal(idx.cpu().int(), torch.from_numpy(g['idx']).int())   # bit-exact vs the reference module's indices


def test_fused_duplicate_codes_take_first_index(V):
    """degenerate codebook: every code duplicated 40 times -> far more exact ties than the candidate list holds (overflow ->
    exact scan of the whole codebook): first-index semantics of torch.argmin, identical to the strict kernel"""
    torch.manual_seed(4)
    N, K, D = 600, 320, 128
    base = torch.randn(8, D)
    cb = base.repeat_interleave(40, dim=0).cuda()
    z = torch.randn(N, D).cuda()
    q0, i0, s0, c0, w0 = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=False)
    q1, i1, s1, c1, w1 = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc='fused')
    assert torch.equal(i0, i1) and torch.equal(q0, q1) and torch.equal(c0, c1)
    assert int((i1 % 40 != 0).sum()) == 0                       # always the first of the 40 identical codes
    assert int(V.ops.vq_assign_raw.last_undecided) == N


@pytest.mark.parametrize('K,n_used', [(8192, 300), (1024, 40), (8192, 8000)])
def test_fused_trained_ema_codebook_large_and_decayed_codes(V, K, n_used):
    """the codebook an EMA run produces (vector_quantizers.py:158-169): a few used codes of the latents' norm among thousands of
    unused ones that have decayed towards the origin.  With one error band for the whole codebook (the largest norm) all decayed
    codes lie within a band of each other and every row overflowed its candidate list (exact scan of all K codes: 30 ms per
    launch in the K = 8192 training step); with per-code bounds the rows are decided from short lists -- and the indices stay
    those of the exact fp32 kernel, for rows next to a used code as for rows next to the origin."""
    torch.manual_seed(12)
    N, D = 4096, 256
    used = torch.randperm(K)[:n_used]
    cb = torch.empty(K, D).uniform_(-1, 1) * 0.027                       # |e| ~ 0.25
    cb[used] = torch.randn(n_used, D)                                    # |e| ~ 16
    z = cb[used[torch.randint(0, n_used, (N,))]] + 0.3 * torch.randn(N, D)
    z[: N // 8] = 0.02 * torch.randn(N // 8, D)                          # latents whose nearest code is one of the decayed ones
    z[N // 8: N // 4] *= 0.05                                            # and latents in between
    z, cb = z.cuda(), cb.cuda()
    q0, i0, s0, c0, w0 = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=False)
    q1, i1, s1, c1, w1 = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc='fused')
    und, full = int(V.ops.vq_assign_raw.last_undecided), int(V.ops.vq_assign_raw.last_fullscan)
    assert torch.equal(i0, i1), (int((i0 != i1).sum()), und, full)
    assert torch.equal(q0, q1) and torch.equal(c0, c1)
    assert full <= 4, (und, full)                                         # no row falls back to the scan of the whole codebook
    assert und <= N // 4, (und, full)


def test_fused_out_of_fp16_range_and_nan_rows(V):
    """the fused search rounds z and the codebook to fp16 for the tensor cores: latents beyond the fp16 range, infinities and
    NaN rows must fall back to the exact scan and still equal the strict kernel (NaN row -> index 0, as the strict kernel)"""
    torch.manual_seed(6)
    N, K, D = 512, 256, 64
    z = torch.randn(N, D)
    z[3, 5] = 1.0e5; z[7, 0] = -3.0e38; z[11, 63] = float('nan'); z[200] *= 7.0e4; z[300, 1] = 6.6e4
    cb = torch.randn(K, D)
    z, cb = z.cuda(), cb.cuda()
    _, i0, _, c0, _ = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=False)
    _, i1, _, c1, _ = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc='fused')
    assert torch.equal(i0, i1) and torch.equal(c0, c1)
    cb2 = cb.clone(); cb2[17, 3] = 1.0e5                              # a codebook entry beyond fp16: every row takes the exact scan
    _, i0, _, _, _ = V.ops.vq_assign_raw(z, cb2, 0, False, False, use_tc=False)
    _, i1, _, _, _ = V.ops.vq_assign_raw(z, cb2, 0, False, False, use_tc='fused')
    assert torch.equal(i0, i1)


def test_fused_prep_cache_follows_the_codebook(V):
    """the quantizer modules cache the bf16 split of their codebook: an in-place change (optimizer step, EMA update, re-init)
    must invalidate it"""
    from vqvae_vqgan_pytorch_lightning_b200.modules.vector_quantizers import EMAVectorQuantizer, VectorQuantizer
    V.set_precision('fast')
    try:
        torch.manual_seed(5)
        x = torch.randn(4, 64, 8, 8).cuda().contiguous(memory_format=torch.channels_last)
        for cls in (VectorQuantizer, EMAVectorQuantizer):
            q = cls(128, 64).cuda().train()
            with torch.no_grad():
                q.codebook.weight.normal_()
                if cls is EMAVectorQuantizer:
                    q.ema_weight.copy_(q.codebook.weight); q.ema_count.fill_(1.0)
            for _ in range(3):
                _, idx, _ = q(x)
                ref = V.ops.vq_assign_raw(x.permute(0, 2, 3, 1).reshape(-1, 64), q.codebook.weight.detach().clone(), 0, False, False,
                                          use_tc=False)[1] if cls is VectorQuantizer else None
                if ref is not None:
                    assert torch.equal(idx.reshape(-1), ref)
                with torch.no_grad():
                    q.codebook.weight.add_(0.3 * torch.randn_like(q.codebook.weight))      # bumps the autograd version
            # EMA: the update kernel rewrites the codebook behind autograd's back and refreshes the cached split itself
            if cls is EMAVectorQuantizer:
                _, idx1, _ = q(x)                                 # runs an EMA update -> new codebook
                cb_now = q.codebook.weight.detach().clone()
                q.eval()
                _, idx2, _ = q(x)
                ref = V.ops.vq_assign_raw(x.permute(0, 2, 3, 1).reshape(-1, 64), cb_now, 0, False, False, use_tc=False)[1]
                assert torch.equal(idx2.reshape(-1), ref)
    finally:
        V.set_precision('strict')
