"""Pins oracle/warp_oracle.py against the reference's outputs (tests/golden/warp.npz, made by
cv2.remap exactly as pdcnet_of.py:34-42 / ofgen.py:37-43 call it) and against cv2 itself."""
import numpy as np
import pytest

from oracle import warp_oracle as wo
from tests import golden_inputs as gi


@pytest.mark.parametrize('name', gi.WARP_CASES)
def test_remap_matches_reference_golden(golden, name):
    img, flow = gi.warp_inputs(name)
    for flavour, fn in (('pdcnet', wo.warp_frame_pdcnet), ('raft', wo.warp_frame_raft)):
        ref = golden['warp'][f'{name}_{flavour}']
        out = fn(img, flow)
        assert out.shape == ref.shape and out.dtype == ref.dtype
        if img.dtype == np.uint8:
            assert np.array_equal(out, ref), f'{name}/{flavour}: {(out != ref).sum()} mismatching bytes'
        else:
            ok = np.isfinite(ref)
            assert np.array_equal(np.isfinite(out), ok)
            np.testing.assert_allclose(out[ok], ref[ok], rtol=0, atol=1e-5 * max(1.0, float(np.abs(img).max())))


def test_remap_exhaustive_fraction_sweep_vs_cv2():
    cv2 = pytest.importorskip('cv2')
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, (16, 16, 3)).astype(np.uint8)
    fy, fx = np.meshgrid(np.arange(32), np.arange(32), indexing='ij')
    mx = (6 + fx / 32).astype(np.float32)
    my = (7 + fy / 32).astype(np.float32)
    ref = cv2.remap(img, mx, my, interpolation=cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
    assert np.array_equal(wo.remap_cubic(img, mx, my), ref)


def test_weight_table_sums_to_one():
    tab = wo.cubic_table_i16().astype(np.int64)
    assert tab.shape == (1024, 16)
    assert (tab.sum(axis=1) == 32768).all()
    # fraction (0,0): the identity tap saturates to int16 (32767) and OpenCV's fix-up puts the missing 1 on tap (2,2)
    assert tab[0, 5] == 32767 and tab[0, 10] == 1 and np.count_nonzero(tab[0]) == 2


def test_bilinear_matches_grid_sample():
    torch = pytest.importorskip('torch')
    F = torch.nn.functional
    img = gi.texture(40, 56, 5).astype(np.float32)
    rs = np.random.RandomState(3)
    flow = (3.0 * rs.standard_normal((40, 56, 2))).astype(np.float32)
    out = wo.warp_bilinear(img, flow)
    H, W = flow.shape[:2]
    xs = torch.arange(W).float()[None, :] + torch.from_numpy(flow[..., 0])
    ys = torch.arange(H).float()[:, None] + torch.from_numpy(flow[..., 1])
    grid = torch.stack([2 * xs / (W - 1) - 1, 2 * ys / (H - 1) - 1], -1)[None]
    ref = F.grid_sample(torch.from_numpy(img).permute(2, 0, 1)[None], grid, mode='bilinear', padding_mode='zeros',
                        align_corners=True)[0].permute(1, 2, 0).numpy()
    # grid_sample's normalise/unnormalise round trip moves the sample point by ~1e-5 px; on this smooth
    # texture (gradient <= ~40/px) that bounds the difference
    np.testing.assert_allclose(out, ref, rtol=0, atol=5e-3)


def test_zero_flow_is_identity():
    img = gi.texture(24, 32, 1)
    flow = np.zeros((24, 32, 2), np.float32)
    assert np.array_equal(wo.warp_frame_pdcnet(img, flow), img)
    assert np.array_equal(wo.warp_frame_raft(img, flow), img)
    assert np.array_equal(wo.warp_bilinear(img, flow), img)


def test_resize_cubic_restatement_matches_cv2():
    """oracle.resize_cubic_f32 (OpenCV's float INTER_CUBIC resize restated) against cv2.resize itself: 3e-7 of the image range
    at the integer ratios warp_frame_latent uses (x8 up, /8 down, pdcnet_of.py:24,30), 3e-6 at arbitrary ratios (OpenCV's
    SIMD / IPP paths fuse and reorder the same float operations)."""
    import cv2
    rs = np.random.RandomState(21)
    for (hs, ws, c, hd, wd, tol) in [(12, 16, 4, 96, 128, 3e-7), (96, 128, 4, 12, 16, 3e-7), (96, 64, 4, 768, 512, 3e-7),
                                      (768, 512, 4, 96, 64, 3e-7), (90, 160, 4, 720, 1280, 3e-7), (17, 23, 3, 50, 41, 3e-6),
                                      (50, 41, 1, 17, 23, 3e-6), (5, 7, 2, 5, 7, 0.0), (33, 20, 4, 100, 7, 3e-6)]:
        img = (3 * rs.standard_normal((hs, ws, c))).astype(np.float32)
        if c == 1:
            img = img[:, :, 0]
        ref = cv2.resize(img, (wd, hd), interpolation=cv2.INTER_CUBIC)
        got = wo.resize_cubic_f32(img, (wd, hd))
        assert got.shape == ref.shape and got.dtype == np.float32
        assert np.abs(got - ref).max() <= tol * np.abs(img).max(), (hs, ws, c, hd, wd)
