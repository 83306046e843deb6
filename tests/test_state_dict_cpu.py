"""CPU: parameter / buffer names, shapes and dtypes of every module on the path equal those of the reference's modules
(fixture: oracle/make_golden_keys.py, produced by constructing the reference classes), so that a reference checkpoint's
state_dict loads with load_state_dict(strict=True) -- SURVEY.md 8(f)-2."""
import json
import os

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def keys():
    with open(os.path.join(HERE, 'golden', 'state_dict_keys.json')) as f:
        return json.load(f)


def describe(m):
    return {k: [list(v.shape), str(v.dtype).replace('torch.', '')] for k, v in m.state_dict().items()}


def build(name):
    from vqvae_vqgan_pytorch_lightning_b200.modules.autoencoder import Encoder, Decoder
    from vqvae_vqgan_pytorch_lightning_b200.modules import vector_quantizers as vq
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.discriminator import Discriminator
    from vqvae_vqgan_pytorch_lightning_b200.modules.loss.lpips import LPIPS
    return {
        'encoder(128,2,[1,2],64)': lambda: Encoder(128, 2, [1, 2], 64),
        'decoder(128,2,[1,2],64)': lambda: Decoder(128, 2, [1, 2], 64),
        'quantizer.standard(32,16)': lambda: vq.VectorQuantizer(32, 16, 0.25),
        'quantizer.ema(32,16)': lambda: vq.EMAVectorQuantizer(32, 16, 0.25, 0.95, 1e-5),
        'quantizer.gumbel(32,16)': lambda: vq.GumbelVectorQuantizer(32, 16, False, 1.0, 5e-4),
        'quantizer.entropy(32,16)': lambda: vq.EntropyVectorQuantizer(32, 16, 0.1, 0.01, 'softmax', 0.25),
        'discriminator(64)': lambda: Discriminator(64),
        'lpips(vgg)': lambda: LPIPS('vgg', pretrained=False),
        'lpips(alex)': lambda: LPIPS('alex', pretrained=False),
    }[name]()


@pytest.mark.parametrize('name', ['encoder(128,2,[1,2],64)', 'decoder(128,2,[1,2],64)', 'quantizer.standard(32,16)',
                                  'quantizer.ema(32,16)', 'quantizer.gumbel(32,16)', 'quantizer.entropy(32,16)',
                                  'discriminator(64)', 'lpips(vgg)', 'lpips(alex)'])
def test_state_dict_layout_matches_reference(keys, name):
    ref = keys[name]
    got = describe(build(name))
    assert list(got.keys()) == list(ref.keys()), (sorted(set(ref) - set(got)), sorted(set(got) - set(ref)))
    for k in ref:
        assert got[k] == ref[k], (k, got[k], ref[k])


def test_reference_style_state_dict_round_trips(keys):
    """a state_dict with the reference's names and shapes loads strictly and the values land in the live parameters."""
    m = build('discriminator(64)')
    sd = {k: torch.full(shape, 0.5, dtype=getattr(torch, dt)) for k, (shape, dt) in keys['discriminator(64)'].items()}
    m.load_state_dict(sd, strict=True)
    assert float(m.b64.conv0.weight.mean()) == 0.5 and float(m.b4.out.bias.mean()) == 0.5
