"""GPU: DevicePrefetcher (pinned host batches -> device one step ahead on a side stream; the reference gets this from
DataLoader(pin_memory=True) + Lightning's batch transfer, vqvae/train.py:121-142) delivers every batch, in order, bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_prefetcher_order_and_content():
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import DevicePrefetcher
    torch.manual_seed(0)
    host = [torch.rand(4, 3, 32, 32).pin_memory() for _ in range(5)]
    dev = torch.device('cuda', 0)
    got = []
    for x in DevicePrefetcher(iter(host), dev):
        assert x.is_cuda
        got.append((x * 1.0).cpu())                       # consume on the current stream
    assert len(got) == len(host)
    for a, b in zip(got, host):
        assert torch.equal(a, b)
    # tuples (image, label) as the reference's loaders yield
    pairs = [(h, i) for i, h in enumerate(host)]
    seen = []
    for xb, label in DevicePrefetcher(iter(pairs), dev).preallocate(pairs[0]):     # (a yielded batch is valid until the next request)
        seen.append(label)
        assert torch.equal(xb.cpu(), host[label])
    assert seen == list(range(5))


@pytest.mark.parametrize('graph', [False, True])
def test_fit_transfers_host_batches(graph):
    """Trainer.fit accepts what the reference's DataLoader yields -- HOST batches, here (images, labels) tuples -- and computes what
    it computes on the same batches already resident on the device."""
    import vqvae_vqgan_pytorch_lightning_b200 as V
    from oracle import init_state as oinit
    from oracle.step_cases import STEP_CASES, q_conf_of
    from tests import common as C
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    V.lib.load()
    V.set_precision('strict')
    case = STEP_CASES['mse_ema']
    torch.manual_seed(11)
    host = [(torch.rand(case['B'], 3, case['S'], case['S']), torch.tensor(i)) for i in range(5)]
    host[1] = (host[1][0].pin_memory(), host[1][1])                     # pinned and pageable batches both work
    out = []
    for batches in (host, [(x.cuda(), y) for x, y in host]):
        sd = oinit.init_state(case['qtype'], case['K'], case['D'], case['ch'], case['nrb'], case['mult'], seed=case['seed'],
                              criterion=None, image_size=case['S'])
        model = V.VQVAE(case['S'], dict(channels=case['ch'], num_res_blocks=case['nrb'], channel_multipliers=list(case['mult'])),
                        q_conf_of(case), None, dict(case['t_conf']), pretrained_lpips=False)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().train()
        model.training_augmentations = None
        loss = Trainer(max_epochs=1, cuda_graph=graph, graph_warmup=2).fit(model, batches)
        assert torch.isfinite(loss).all()
        out.append({k: v.detach().clone() for k, v in model.state_dict().items()})
    num = den = 0.0
    for k in out[0]:
        if k in C.DEGENERATE or not out[0][k].dtype.is_floating_point:
            continue
        num += float((out[0][k].double() - out[1][k].double()).pow(2).sum()); den += float(out[1][k].double().pow(2).sum())
    assert (num / den) ** 0.5 < 1e-3, (num / den) ** 0.5              # two runs differ by the weight-gradient atomics only
