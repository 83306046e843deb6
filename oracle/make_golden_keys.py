"""Generate tests/golden/state_dict_keys.json by CONSTRUCTING THE REFERENCE'S OWN modules (test infrastructure only):
parameter / buffer names, shapes and dtypes of every module on the hot path, i.e. what a reference Lightning checkpoint's
`state_dict` holds under `encoder.`, `decoder.`, `quantizer.`, `criterion.` (vqvae/model.py:79-149).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_keys.py
"""
from __future__ import annotations

import json
import os
import sys
from collections import OrderedDict

import torch

REF = os.environ.get('VQ_REF_PATH', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'state_dict_keys.json')


def describe(m: torch.nn.Module):
    return OrderedDict((k, [list(v.shape), str(v.dtype).replace('torch.', '')]) for k, v in m.state_dict().items())


def main():
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import torchvision
    from vqvae.modules.autoencoder import Encoder, Decoder
    from vqvae.modules import vector_quantizers as vq
    from vqvae.modules.loss.stylegan2_discriminator.discriminator import Discriminator
    import vqvae.modules.loss.lpips_pytorch.modules.networks as nets
    import vqvae.modules.loss.lpips_pytorch.modules.lpips as lp
    _vgg = torchvision.models.vgg16
    nets.models.vgg16 = lambda weights=None, **kw: _vgg(weights=None)
    lp.get_state_dict = lambda net_type='alex', version='0.1': OrderedDict(
        (f'{i}.1.weight', torch.rand(1, c, 1, 1)) for i, c in enumerate([64, 128, 256, 512, 512]))
    out = OrderedDict()
    out['encoder(128,2,[1,2],64)'] = describe(Encoder(128, 2, [1, 2], 64))
    out['decoder(128,2,[1,2],64)'] = describe(Decoder(128, 2, [1, 2], 64))
    out['quantizer.standard(32,16)'] = describe(vq.VectorQuantizer(32, 16, 0.25))
    out['quantizer.ema(32,16)'] = describe(vq.EMAVectorQuantizer(32, 16, 0.25, 0.95, 1e-5))
    out['quantizer.gumbel(32,16)'] = describe(vq.GumbelVectorQuantizer(32, 16, False, 1.0, 5e-4))
    out['quantizer.entropy(32,16)'] = describe(vq.EntropyVectorQuantizer(32, 16, 0.1, 0.01, 'softmax', 0.25))
    out['discriminator(64)'] = describe(Discriminator(64))
    out['lpips(vgg)'] = describe(lp.LPIPS('vgg'))
    _alex = torchvision.models.alexnet
    nets.models.alexnet = lambda *a, **kw: _alex(weights=None)
    lp.get_state_dict = lambda net_type='alex', version='0.1': OrderedDict(
        (f'{i}.1.weight', torch.rand(1, c, 1, 1)) for i, c in enumerate([64, 192, 384, 256, 256]))
    out['lpips(alex)'] = describe(lp.LPIPS('alex'))
    with open(OUT, 'w') as f:
        json.dump(out, f, indent=0)
    print({k: len(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
