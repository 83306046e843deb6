"""CPU, world_size 2, gloo: the host side of the data-parallel path -- one SUM all-reduce of the flat gradient buffer,
averaging folded into the optimizer's grad_scale, and the EMA cluster-statistics all-reduce that makes N ranks equal
the single-process reference on the concatenated batch.  Kernels are not run here (no GPU): rank-local
statistics come from the oracle, the exchange goes through the product's Trainer hooks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vqvae_oracle as orc


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


class _FakeOpt:
    def __init__(self, n):
        self.flat_grad = torch.zeros(n)
        self.grad_scale = 1.0
        self.param_groups = []


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    tr = Trainer()
    assert tr.world_size == world and tr.rank == rank
    # --- gradient path: one all-reduce of the flat buffer, mean folded into grad_scale
    opt = _FakeOpt(1000)
    tr.optimizers = [opt]
    opt.grad_scale = 1.0 / tr.world_size
    torch.manual_seed(100 + rank)
    local = torch.randn(1000)
    opt.flat_grad.copy_(local)
    tr.sync_gradients()
    all_local = [torch.empty(1000) for _ in range(world)]
    dist.all_gather(all_local, local)
    assert torch.allclose(opt.flat_grad * opt.grad_scale, torch.stack(all_local).mean(0), atol=1e-6)
    # --- EMA statistics: shard a global batch, all-reduce (counts, dw), update with the GLOBAL image batch
    torch.manual_seed(7)
    K, D, B = 32, 16, 8
    cb = torch.randn(K, D); ema_w = cb.clone(); ema_c = torch.zeros(K)
    z = torch.randn(B, D, 4, 4)
    ref = orc.vq_ema(z, cb, ema_c, ema_w, 0.25, 0.95, 1e-5, True)                  # single process, whole batch
    zl = z[rank * (B // world):(rank + 1) * (B // world)]
    flat = zl.permute(0, 2, 3, 1).reshape(-1, D)
    idx = torch.argmin(orc.l2_distances(flat, cb), dim=1)
    counts = torch.bincount(idx, minlength=K).float()
    dw = torch.zeros(K, D).index_add_(0, idx, flat)
    tr._allreduce_stats(counts, dw)
    cnt = ema_c * 0.95 + 0.05 * counts
    b = zl.shape[0] * world
    new_cnt = (cnt + 1e-5) / (b + K * 1e-5) * b
    new_w = ema_w * 0.95 + 0.05 * dw
    assert torch.allclose(new_cnt, ref[4], atol=1e-6) and torch.allclose(new_w, ref[5], atol=1e-5)
    assert torch.allclose(new_w / new_cnt[:, None], ref[3], rtol=1e-4, atol=1e-5)
    if rank == 0:
        out.put('ok')
    dist.destroy_process_group()


def test_two_rank_gradient_and_ema_exchange():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 'ok'


def test_lr_and_batch_derivation():
    from vqvae_vqgan_pytorch_lightning_b200.common_utils import derive_confs, get_model_conf
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    conf = get_model_conf(os.path.join(root, 'example_confs', 'ema_vqvae.yaml'))
    image_size, ae, q, l, t, bs = derive_confs(conf, world_size=8)
    assert image_size == 256 and bs == 32 and l is None and abs(t['lr'] - 1e-4) < 1e-12        # train.py:59-63
    assert q['num_embeddings'] == 4096 and q['params']['epsilon'] == '1e-5'                     # YAML string quirk kept
    _, _, q2, _, t2, bs2 = derive_confs(conf, 1, {'num_embeddings': 1024, 'cumulative_bs': 64})
    assert q2['num_embeddings'] == 1024 and bs2 == 64 and abs(t2['lr'] - 0.5e-4) < 1e-12
