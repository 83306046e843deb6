"""CPU: the crop-box / flip sampler of the fused training augmentation (augment.py) -- integer, square, in-bounds boxes whose
area follows scale ~ U(0.7, 1), per-sample flips with p = 0.5 (the documented behaviour of kornia's RandomResizedCrop(scale=(0.7, 1),
ratio=(1, 1)) + RandomHorizontalFlip used by the reference, base_autoencoder.py:17-22)."""
import torch

from vqvae_vqgan_pytorch_lightning_b200.augment import RandomResizedCropFlip


def test_sampler_boxes_are_integer_square_in_bounds():
    g = torch.Generator().manual_seed(0)
    aug = RandomResizedCropFlip(256)
    boxes, flip = aug.sample(4096, 256, 256, 'cpu', generator=g)
    assert boxes.shape == (4096, 4) and flip.shape == (4096,) and flip.dtype == torch.uint8
    assert torch.equal(boxes, boxes.round())
    side = boxes[:, 2] - boxes[:, 0] + 1
    assert torch.equal(side, boxes[:, 3] - boxes[:, 1] + 1)
    assert (boxes[:, :2] >= 0).all() and (boxes[:, 2:] <= 255).all()
    area = (side * side) / (256.0 * 256.0)
    assert float(area.min()) >= 0.69 and float(area.max()) <= 1.0
    assert abs(float(area.mean()) - 0.85) < 0.01                     # U(0.7, 1) has mean 0.85
    assert abs(float(flip.float().mean()) - 0.5) < 0.03
    # corners are uniform over the admissible range: a full-size crop can only start at 0
    full = side == 256
    assert (boxes[full][:, :2] == 0).all()


def test_sampler_is_per_sample_and_rectangular_sources_are_clamped():
    g = torch.Generator().manual_seed(1)
    aug = RandomResizedCropFlip(64)
    boxes, _ = aug.sample(512, 48, 80, 'cpu', generator=g)
    side = boxes[:, 2] - boxes[:, 0] + 1
    # sqrt(U(0.7, 1) * 48 * 80) >= 51.8 > 48: every crop is clamped to the short edge, only the x offset (0..32) varies
    assert float(side.min()) == 48 and float(side.max()) == 48 and len(torch.unique(boxes, dim=0)) >= 30      # same_on_batch=False
    assert (boxes[:, 3] <= 47).all() and (boxes[:, 2] <= 79).all()
