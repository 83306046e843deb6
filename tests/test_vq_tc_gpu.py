"""GPU: the tensor-core nearest-code search (bf16 hi/lo split + exact fp32 re-evaluation of near-ties) must return the
same indices as the exact fp32 kernel -- bit-exact for tie-free codebooks, and via the exact fallback for the reference's
tie-heavy U(+-1/K) initial codebook."""
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


@pytest.mark.parametrize('N,K,D,init', [(4096, 1024, 256, 'normal'), (4096, 1024, 256, 'uniform'), (1000, 512, 64, 'normal'),
                                        (16384, 1024, 256, 'normal'), (2048, 8192, 256, 'normal'), (300, 264, 128, 'trained')])
def test_tc_search_equals_exact_kernel(V, N, K, D, init):
    torch.manual_seed(9)
    z = torch.randn(N, D).cuda()
    if init == 'uniform':
        cb = torch.empty(K, D).uniform_(-1 / K, 1 / K).cuda()
    elif init == 'trained':
        cb = (z[torch.randint(0, N, (K,))] + 0.05 * torch.randn(K, D, device='cuda'))      # codes sit near the data
    else:
        cb = torch.randn(K, D).cuda()
    q0, i0, s0, c0, w0 = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=False)
    q1, i1, s1, c1, w1 = V.ops.vq_assign_raw(z, cb, 0, True, True, use_tc=True)
    und = int(V.ops.vq_assign_raw.last_undecided)
    assert torch.equal(i0, i1), (int((i0 != i1).sum()), und)
    assert torch.equal(q0, q1)
    assert abs(float(s0) - float(s1)) <= 1e-6 * abs(float(s0))
    assert torch.equal(c0, c1) and C.rel_err(w1, w0) < 1e-5
    if init == 'normal':
        assert und <= N // 20, und                 # tie-free: almost every row is decided by the tensor-core pass
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
    for order in (0, 1):
        _, idx, _, _, _ = V.ops.vq_assign_raw(flat, cb.cuda(), order, False, False, use_tc=True)
        assert torch.equal(idx.cpu().int(), torch.from_numpy(g['idx']).int())       # bit-exact vs the reference module's indices
