"""Pins oracle/blur_oracle.py (the NumPy restatement of Pillow's GaussianBlur / composite / BICUBIC resize, the
arithmetic behind guided_ldm_inpainting.py:290-309) against Pillow itself -- the library the reference calls."""
import numpy as np
import pytest
from PIL import Image, ImageFilter

from oracle import blur_oracle as bo


@pytest.mark.parametrize('hw', [(64, 80), (7, 5), (33, 130), (1, 9)])
@pytest.mark.parametrize('radius', [0, 0.1, 0.5, 1, 2, 4, 5.5, 8, 12])
def test_gaussian_blur_bit_exact_vs_pillow(hw, radius):
    rs = np.random.RandomState(int(radius * 10) + hw[0])
    for m in ((rs.rand(*hw) < 0.3).astype(np.uint8) * 255, rs.randint(0, 256, hw).astype(np.uint8)):
        ref = np.array(Image.fromarray(m).convert('L').filter(ImageFilter.GaussianBlur(radius)))
        assert np.array_equal(bo.gaussian_blur_u8(m, radius), ref)


def test_composite_bit_exact_vs_pillow():
    rs = np.random.RandomState(1)
    a = rs.randint(0, 256, (40, 50, 3)).astype(np.uint8)
    b = rs.randint(0, 256, (40, 50, 3)).astype(np.uint8)
    m = rs.randint(0, 256, (40, 50)).astype(np.uint8)
    m[:4] = 0
    m[4:8] = 255
    ref = np.array(Image.composite(Image.fromarray(a), Image.fromarray(b), Image.fromarray(m)))
    out = bo.composite(a, b, m)
    assert np.array_equal(out, ref)
    assert np.array_equal(out[:4], b[:4]) and np.array_equal(out[4:8], a[4:8])


@pytest.mark.parametrize('shape', [(192, 128, 24, 16), (64, 80, 8, 10), (100, 60, 12, 7), (16, 16, 16, 16), (40, 40, 80, 60), (9, 9, 1, 1)])
def test_resize_bicubic_bit_exact_vs_pillow(shape):
    H, W, oh, ow = shape
    m = np.random.RandomState(H).randint(0, 256, (H, W)).astype(np.uint8)
    ref = np.array(Image.fromarray(m).convert('RGB').resize((ow, oh)))[:, :, 0]
    assert np.array_equal(bo.resize_bicubic_u8(m, ow, oh), ref)


def test_reference_step_end_to_end_vs_pillow():
    """The literal expressions of guided_ldm_inpainting.py:290-308 (with reference_img) against the oracle."""
    rs = np.random.RandomState(3)
    H, W = 96, 64
    image = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    reference = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    mask = np.zeros((H, W), np.uint8)
    mask[20:50, 10:40] = 255
    mask[70:, 50:] = 255
    image_mask = Image.fromarray(mask).convert('L').filter(ImageFilter.GaussianBlur(4))
    comp = np.array(Image.composite(Image.fromarray(reference), Image.fromarray(image), image_mask))
    latmask = image_mask.convert('RGB').resize((W // 8, H // 8))
    latmask = np.moveaxis(np.array(latmask, dtype=np.float32), 2, 0) / 255
    latmask = np.tile(np.around(latmask[0])[None], (4, 1, 1))
    out, blurred, lat = bo.blur_composite_latmask(image, reference, mask, 4)
    assert np.array_equal(blurred, np.array(image_mask))
    assert np.array_equal(out, comp)
    assert np.array_equal(lat, latmask)
