import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vqvae_vqgan_pytorch_lightning_b200 as pkg
from tests import common as C
pkg.lib.load(); pkg.set_precision('fast')
sd, x = C.seeded_inputs('cfg1', 'standard'); c = C.CASES['cfg1']
model = pkg.VQVAE(c['S'], dict(channels=c['ch'], num_res_blocks=c['nrb'], channel_multipliers=list(c['mult'])),
                  dict(num_embeddings=c['K'], embedding_dim=c['D'], type='standard', params=dict(commitment_cost=0.25), reinit_every_n_epochs=None),
                  None, dict(lr=1e-4, betas=[0.0, 0.99], eps=1e-8, weight_decay=1e-4, warmup_epochs=None, decay_epochs=None))
model.load_state_dict(sd); model = model.cuda().train()
def hook(name):
    def f(m, i, o):
        t = o[0] if isinstance(o, tuple) else o
        if torch.is_tensor(t) and t.is_floating_point():
            bad = int((~torch.isfinite(t.float())).sum())
            if bad: print('NONFINITE', name, tuple(t.shape), t.dtype, bad, flush=True)
    return f
for n, m in model.named_modules():
    if n: m.register_forward_hook(hook(n))
xg = x.cuda().contiguous(memory_format=torch.channels_last)
y = model.encoder.conv_in(xg)
print('conv_in out finite:', bool(torch.isfinite(y.float()).all()), y.dtype, float(y.float().abs().max()))
recon, ql, idx = model(xg)
print('done', float(ql))
