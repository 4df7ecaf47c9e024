"""GPU parity of the after-the-path step (mask blur, composite, latent mask) through the C ABI:
bit-exact against the oracle (itself pinned against Pillow) and against Pillow directly."""
import numpy as np
import pytest
import torch
from PIL import Image, ImageFilter

from oracle import blur_oracle as bo

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize('hw', [(768, 512), (96, 64), (61, 45), (7, 5), (300, 700)])
@pytest.mark.parametrize('radius', [0, 0.5, 2, 4, 9.5, 30])
def test_gaussian_blur_bit_exact(cuda, hw, radius):
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(hw[0] + int(radius))
    B = 2
    m = np.stack([(rs.rand(*hw) < 0.3).astype(np.uint8) * 255, rs.randint(0, 256, hw).astype(np.uint8)])
    _, blurred = ops.mask_blur_composite(_t(m, cuda), None, None, radius)
    got = blurred.cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b], bo.gaussian_blur_u8(m[b], radius)), f'image {b}: {(got[b] != bo.gaussian_blur_u8(m[b], radius)).sum()} bytes differ'
    if hw[0] <= 96:
        assert np.array_equal(got[0], np.array(Image.fromarray(m[0]).filter(ImageFilter.GaussianBlur(radius))))


def test_blur_composite_and_latmask_vs_pillow(cuda):
    from sd_animation_optical_flow_b200 import guided_ldm_inpainting as gli
    rs = np.random.RandomState(5)
    H, W = 768, 512
    image = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    reference = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    mask = (rs.rand(H // 16, W // 16) < 0.3).astype(np.uint8).repeat(16, 0).repeat(16, 1) * 255
    image_t, image_mask, nmask = gli.prepare_inpaint_inputs(Image.fromarray(image), Image.fromarray(mask), 4, Image.fromarray(reference))
    # the reference's literal expressions (guided_ldm_inpainting.py:290-308)
    pm = Image.fromarray(mask).convert('L').filter(ImageFilter.GaussianBlur(4))
    comp = np.array(Image.composite(Image.fromarray(reference), Image.fromarray(image), pm)).astype(np.float32) / 127.5 - 1.0
    lat = np.moveaxis(np.array(pm.convert('RGB').resize((W // 8, H // 8)), dtype=np.float32), 2, 0) / 255
    lat = np.tile(np.around(lat[0])[None], (4, 1, 1))
    assert np.array_equal(image_mask, np.array(pm))
    assert image_t.shape == (1, 3, H, W) and np.array_equal(image_t[0].cpu().numpy(), np.moveaxis(comp, 2, 0))
    assert nmask.shape == (4, H // 8, W // 8) and np.array_equal(nmask.cpu().numpy(), lat)
    # RGB (grey) masks are accepted like PIL's convert('L')
    _, im2, _ = gli.prepare_inpaint_inputs(image, np.repeat(mask[:, :, None], 3, 2), 4, reference)
    assert np.array_equal(im2, image_mask)
    with pytest.raises(NotImplementedError):
        gli.prepare_inpaint_inputs(image, mask, 4, None)


@pytest.mark.parametrize('shape', [(768, 512, 96, 64), (720, 1280, 90, 160), (100, 60, 12, 7), (40, 40, 80, 60), (16, 16, 16, 16)])
def test_resize_bicubic_bit_exact(cuda, shape):
    from sd_animation_optical_flow_b200 import ops
    H, W, oh, ow = shape
    m = np.random.RandomState(W).randint(0, 256, (2, H, W)).astype(np.uint8)
    dst, lat = ops.resize_bicubic_u8(_t(m, cuda), oh, ow, want_latmask=True)
    for b in range(2):
        ref = bo.resize_bicubic_u8(m[b], ow, oh)
        assert np.array_equal(dst[b].cpu().numpy(), ref)
        assert np.array_equal(lat[b].cpu().numpy(), np.tile(np.around(ref.astype(np.float32) / 255)[None], (4, 1, 1)))
