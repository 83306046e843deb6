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
