"""CPU: host-side logic that needs no kernel -- schedules (product vs the oracle's restatement and the call-site semantics of
vqvae/model.py:163-230), the Lightning `.ckpt` loader (vqvae/train.py:106-111, evaluate.py:48-49), and the data-parallel
code re-initialisation (base_quantizer.py:81-102 with summed usage counts and a broadcast draw; gloo, world size 2)."""
import json
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vqvae_oracle as orc
from tests import common as C
from vqvae_vqgan_pytorch_lightning_b200 import schedulers as S


def test_schedulers_match_oracle_and_call_site_semantics():
    lin, cos = S.LinearScheduler(0, 100, 1e-20, 1e-3), S.CosineScheduler(0, 400, 1e-3, 5e-4)
    lc = S.LinearCosineScheduler(0, 400, 1e-3, 5e-4, 100)          # model.py:175: warm-up to lr until th_step, cosine to lr/2
    for i in (0, 1, 37, 99, 100, 101, 250, 399, 400, 1000):
        assert lin.step(i) == orc.linear_schedule(i, 0, 100, 1e-20, 1e-3)
        assert cos.step(i) == orc.cosine_schedule(i, 0, 400, 1e-3, 5e-4)
        assert lc.step(i) == orc.linear_cosine_schedule(i, 0, 400, 1e-3, 5e-4, 100)
    assert lin.step(0) == 1e-20 and lin.step(100) == 1e-3 and abs(lin.step(50) - 5e-4) < 1e-12
    assert cos.step(0) == 1e-3 and cos.step(400) == 5e-4 and abs(cos.step(200) - 7.5e-4) < 1e-12
    vals = [lc.step(i) for i in range(0, 401)]
    assert all(b >= a for a, b in zip(vals[:100], vals[1:101]))          # monotone warm-up
    assert all(b <= a for a, b in zip(vals[100:400], vals[101:401]))      # monotone decay
    assert abs(vals[100] - 1e-3) < 1e-12 and abs(vals[400] - 5e-4) < 1e-15
    kl = S.CosineScheduler(0, int(0.48 * 1000), 0.0, 0.00859375)         # gumbel kl warm-up (model.py:190-195)
    assert kl.step(0) == 0.0 and kl.step(480) == 0.00859375 and 0 < kl.step(240) < 0.00859375
    lin.destroy(); cos.destroy(); lc.destroy()                            # model.py:305-307 calls destroy()


def test_load_from_checkpoint_reference_layout():
    """a file laid out like the reference's Lightning checkpoints: the state_dict of a reference VQVAE instance (names checked
    against the fixture written by executing the reference class, tensors from the seeded construction that equals it)"""
    from oracle import init_state as oinit
    from oracle.step_cases import STEP_CASES, q_conf_of
    from vqvae_vqgan_pytorch_lightning_b200.model import VQVAE
    case = STEP_CASES['gan_hinge_adaptive_r1']
    g = C.golden('step_gan_hinge_adaptive_r1')
    sd = oinit.init_state(case['qtype'], case['K'], case['D'], case['ch'], case['nrb'], case['mult'], seed=case['seed'],
                          criterion='gan', image_size=case['S'])
    assert set(sd) == set(g['init_names'].tolist())
    ae_conf = dict(channels=case['ch'], num_res_blocks=case['nrb'], channel_multipliers=list(case['mult']))
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'last.ckpt')
        torch.save({'state_dict': sd, 'epoch': 7, 'global_step': 1234, 'pytorch-lightning_version': '2.0.0'}, path)
        # the reference's inference call (evaluate.py:48-49): strict=False, init_cb=False, load_loss=False -> criterion.* ignored
        m = VQVAE.load_from_checkpoint(path, strict=False, image_size=case['S'], ae_conf=ae_conf, q_conf=q_conf_of(case), l_conf=None,
                                       t_conf=None, init_cb=False, load_loss=False)
        own = m.state_dict()
        assert set(own) == {k for k in sd if not k.startswith('criterion.')}
        for k, v in own.items():
            assert torch.equal(v, sd[k]), k
        assert m._checkpoint_extras['epoch'] == 7 and m._checkpoint_extras['global_step'] == 1234
        with pytest.raises(RuntimeError):                                   # strict=True must complain about the criterion.* keys
            VQVAE.load_from_checkpoint(path, strict=True, image_size=case['S'], ae_conf=ae_conf, q_conf=q_conf_of(case), l_conf=None,
                                       t_conf=None, init_cb=False, load_loss=False)
        # the training-resume call (train.py:106-111): the whole model incl. the loss heads, strict
        m2 = VQVAE.load_from_checkpoint(path, strict=True, image_size=case['S'], ae_conf=ae_conf, q_conf=q_conf_of(case),
                                        l_conf=case['l_conf'], t_conf=dict(case['t_conf']), init_cb=False, pretrained_lpips=False)
        own = m2.state_dict()
        assert set(own) == set(sd)
        for k, v in own.items():
            assert torch.equal(v, sd[k]), k


def _reinit_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from vqvae_vqgan_pytorch_lightning_b200.modules.vector_quantizers import VectorQuantizer
    torch.manual_seed(3)                                                      # identical replicas
    q = VectorQuantizer(32, 8)
    q.init_codebook()
    torch.manual_seed(100 + rank)                                             # per-rank RNG streams, as in a real DP run
    counts = torch.zeros(32)
    counts[(rank * 5):(rank * 5 + 6)] = torch.arange(1, 7).float()           # each rank saw different codes
    usage = q.get_codebook_usage(q.reduce_usage(counts))[0]
    q.reinit_unused_codes(usage)
    out[rank] = (q.codebook.weight.detach().clone(), usage.clone())
    dist.destroy_process_group()


def test_reinit_unused_codes_keeps_replicas_identical():
    world, port = 2, 29500 + os.getpid() % 500
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_reinit_worker, args=(world, port, out), nprocs=world, join=True)
    cb0, u0 = out[0]; cb1, u1 = out[1]
    assert torch.equal(u0, u1) and torch.equal(cb0, cb1)                      # same global usage, same draw on every rank
    total = torch.zeros(32); total[0:6] += torch.arange(1, 7).float(); total[5:11] += torch.arange(1, 7).float()
    assert torch.allclose(u0, total / total.sum())
    unused = (u0 == 0).nonzero().flatten()
    live = (u0 > 0).nonzero().flatten()
    torch.manual_seed(3)
    from vqvae_vqgan_pytorch_lightning_b200.modules.vector_quantizers import VectorQuantizer
    ref = VectorQuantizer(32, 8); ref.init_codebook()
    for i in unused.tolist():                                                 # every dead code now equals some live code
        assert any(torch.equal(cb0[i], ref.codebook.weight[j]) for j in live.tolist())
    for j in live.tolist():
        assert torch.equal(cb0[j], ref.codebook.weight[j])


def test_trainer_batch_transfer_passes_through_without_a_device():
    """Trainer._device_batches: only HOST batches of a model that lives on a GPU go through the DevicePrefetcher; anything else
    (here: a CPU module) is handed on unchanged and in order, tuples included, and an empty loader stays empty."""
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    m = torch.nn.Linear(2, 2)
    batches = [(torch.full((1, 2), float(i)), torch.tensor(i)) for i in range(4)]
    got = list(Trainer._device_batches(batches, m))
    assert len(got) == 4 and all(a is b for a, b in zip(got, batches))
    assert list(Trainer._device_batches([], m)) == []
    assert list(Trainer._device_batches(iter(batches), m))[3][1].item() == 3          # one-shot iterators work too
