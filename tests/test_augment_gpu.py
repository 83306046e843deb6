"""GPU: the fused input pipeline (clamp -> resized crop -> flip -> normalise, vqb_crop_flip_normalize) against torch's own
crop + F.interpolate(bilinear, align_corners=True) + flip + normalise on the SAME boxes (the semantics of the reference's
kornia chain, base_autoencoder.py:17-50; the random parameter stream itself is parity-unpinned, see augment.py)."""
import pytest
import torch
import torch.nn.functional as F

from tests import common as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def V():
    import vqvae_vqgan_pytorch_lightning_b200 as pkg
    pkg.lib.load()
    pkg.set_precision('strict')
    return pkg


def torch_pipeline(images01, boxes, flip, S):
    out = []
    for n in range(images01.shape[0]):
        x0, y0, x1, y1 = [int(v) for v in boxes[n].tolist()]
        crop = images01[n:n + 1, :, y0:y1 + 1, x0:x1 + 1].clamp(0, 1)
        r = F.interpolate(crop, size=(S, S), mode='bilinear', align_corners=True)
        if flip[n]:
            r = r.flip(-1)
        out.append((r - 0.5) / 0.5)
    return torch.cat(out)


@pytest.mark.parametrize('in_dtype', [torch.float32, torch.float16, torch.uint8])
def test_crop_flip_normalize_matches_torch(V, in_dtype):
    from vqvae_vqgan_pytorch_lightning_b200.augment import RandomResizedCropFlip, crop_flip_normalize
    torch.manual_seed(0)
    n, c, h, w, S = 6, 3, 40, 52, 32
    if in_dtype == torch.uint8:
        src = torch.randint(0, 256, (n, c, h, w), dtype=torch.uint8)
        ref_in = src.float() / 255.0
    else:
        src = (torch.rand(n, c, h, w) * 1.2 - 0.1).to(in_dtype)          # a few values outside [0,1]: the clamp matters
        ref_in = src.float()
    aug = RandomResizedCropFlip(S)
    boxes, flip = aug.sample(n, h, w, 'cuda')
    b = boxes.cpu()
    side = b[:, 2] - b[:, 0] + 1
    assert torch.equal(b, b.round()) and torch.equal(side, b[:, 3] - b[:, 1] + 1)             # integer, square boxes
    assert (b[:, 0] >= 0).all() and (b[:, 1] >= 0).all() and (b[:, 2] <= w - 1).all() and (b[:, 3] <= h - 1).all()
    assert (side >= 38).all() and (side <= 40).all()              # round(sqrt(U(0.7,1) * 40 * 52)) clamped to the short edge
    got = crop_flip_normalize(src.cuda(), boxes, flip, (S, S))
    assert got.shape == (n, c, S, S) and got.is_contiguous(memory_format=torch.channels_last)
    ref = torch_pipeline(ref_in, b, flip.cpu(), S)
    assert float((got.cpu() - ref).abs().max()) < 2e-5
    got16 = crop_flip_normalize(src.cuda(), boxes, flip, (S, S), out_dtype=torch.bfloat16)
    assert float((got16.float().cpu() - ref).abs().max()) < 8e-3


def test_identity_box_equals_plain_preprocessing(V):
    """the full-image box without flip reproduces clamp + normalise exactly (the validation-time preprocess_batch)."""
    from vqvae_vqgan_pytorch_lightning_b200.augment import crop_flip_normalize
    torch.manual_seed(1)
    x = torch.rand(3, 3, 24, 24).cuda()
    boxes = torch.tensor([[0, 0, 23, 23]] * 3, dtype=torch.float32)
    got = crop_flip_normalize(x, boxes, None, (24, 24))
    ref = V.ops.images_to_nhwc(x, torch.float32, normalize=True)
    assert float((got - ref).abs().max()) < 1e-6


def test_training_step_runs_with_default_augmentation(V):
    """the reference's default: augmentation ON in training_step, OFF in validation / inference preprocessing."""
    from vqvae_vqgan_pytorch_lightning_b200.lightning_shim import Trainer
    torch.manual_seed(2)
    model = V.VQVAE(32, dict(channels=32, num_res_blocks=1, channel_multipliers=[1, 2]),
                    dict(num_embeddings=32, embedding_dim=32, type='standard', params=dict(commitment_cost=0.25), reinit_every_n_epochs=None),
                    None, dict(lr=1e-4, betas=[0.9, 0.99], eps=1e-8, weight_decay=0.0, warmup_epochs=None, decay_epochs=None)).cuda().train()
    assert getattr(model.training_augmentations, 'fused', False)
    tr = Trainer(max_epochs=1, num_training_batches=2)
    tr.attach(model); model.on_train_start()
    x = torch.rand(4, 3, 32, 32, device='cuda')
    a = model.preprocess_batch(x, training=True); b = model.preprocess_batch(x, training=False)
    assert a.shape == b.shape and not torch.equal(a, b)
    assert torch.equal(b, V.ops.images_to_nhwc(x, torch.float32, normalize=True))
    loss = tr.run_step(x, 0)
    assert torch.isfinite(loss).all()
    u8 = (x * 255).round().to(torch.uint8)
    assert float((model.preprocess_batch(u8, training=False) - V.ops.images_to_nhwc(u8.float() / 255, torch.float32, True)).abs().max()) < 1e-6
