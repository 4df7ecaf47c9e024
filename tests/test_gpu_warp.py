"""GPU parity of the warp kernels, through the C ABI (ctypes) and the reference-named numpy API.
Bar: u8 cubic bit-exact vs the reference's cv2.remap outputs (golden) and vs the oracle at full
size; float cubic 1e-5 relative to the image range; bilinear 1e-4 abs on [0,255] vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import warp_oracle as wo
from tests import golden_inputs as gi

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize('name', gi.WARP_CASES)
def test_numpy_api_matches_reference_golden(cuda, golden, name):
    from sd_animation_optical_flow_b200 import ofgen, pdcnet_of
    img, flow = gi.warp_inputs(name)
    flow_before = flow.copy()
    for flavour, fn in (('pdcnet', pdcnet_of.warp_frame), ('raft', ofgen.warp_frame)):
        ref = golden['warp'][f'{name}_{flavour}']
        out = fn(img, flow)
        assert out.shape == ref.shape and out.dtype == ref.dtype
        if img.dtype == np.uint8:
            assert np.array_equal(out, ref), f'{name}/{flavour}: {(out != ref).sum()} of {ref.size} bytes differ'
        else:
            ok = np.isfinite(ref)
            np.testing.assert_allclose(out[ok], ref[ok], rtol=0, atol=1e-5 * max(1.0, float(np.abs(img).max())))
    assert np.array_equal(flow, flow_before, equal_nan=True), 'warp_frame must not modify flow'


@pytest.mark.parametrize('hw', [(768, 512), (720, 1280), (61, 45)])
def test_cubic_u8_full_size_bit_exact_vs_oracle(cuda, hw):
    from sd_animation_optical_flow_b200 import ops
    H, W = hw
    rs = np.random.RandomState(H)
    img = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    flow = (12.0 * rs.standard_normal((H, W, 2))).astype(np.float32)
    flow[::7, ::5] *= 40.0   # a sprinkling of far out-of-image samples
    for sign, fn in ((1.0, wo.warp_frame_pdcnet), (-1.0, wo.warp_frame_raft)):
        out = ops.warp(_t(img, cuda), _t(flow, cuda), 'cv2_cubic', sign).cpu().numpy()
        ref = fn(img, flow)
        assert np.array_equal(out, ref), f'{(out != ref).sum()} bytes differ'


def smooth_flow(B, H, W, seed, dev):
    """The flow generator of bench.py's warp roofline (camera / object motion): per frame a random translation of a few
    pixels plus a low-frequency deformation (1/64-resolution Gaussian field of 4 px, bicubic-upsampled: |grad| ~ 0.1)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    lo = torch.randn((B, 2, -(-H // 64), -(-W // 64)), generator=g, device=dev) * 4
    f = torch.nn.functional.interpolate(lo, scale_factor=64, mode='bicubic', align_corners=False)[:, :, :H, :W]
    return (f + 6 * torch.randn((B, 2, 1, 1), generator=g, device=dev)).permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize('hw', [(768, 512), (720, 1280)])
def test_cubic_u8_full_size_smooth_flow_takes_the_staged_path(cuda, hw):
    """The STAGED shared-memory path of the tiled kernel (the one the warp roofline is quoted on) at full size, bit-exact
    against the oracle, with the kernel's own tile counters proving which path ran (the noisy-flow
    full-size test above only exercises the per-pixel fallback)."""
    from sd_animation_optical_flow_b200 import ops
    H, W = hw
    B = 3
    rs = np.random.RandomState(W)
    imgs = rs.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    flow = smooth_flow(B, H, W, 11, cuda)
    ops.warp_tile_stats(reset=True)
    out = ops.warp(_t(imgs, cuda), flow, 'cv2_cubic', 1.0).cpu().numpy()
    staged, fallback = ops.warp_tile_stats(reset=True)
    ntiles = B * -(-H // 32) * -(-W // 32)
    assert staged + fallback == ntiles
    print(f'{H}x{W}: {staged} of {ntiles} tiles staged ({100.0 * staged / ntiles:.1f} %)')
    assert staged >= 0.95 * ntiles
    fl = flow.cpu().numpy()
    for b in range(B):
        ref = wo.warp_frame_pdcnet(imgs[b], fl[b])
        assert np.array_equal(out[b], ref), f'frame {b}: {(out[b] != ref).sum()} bytes differ'
    # the same frames warped with x - flow (ofgen.warp_frame) and a shared source (one key frame, B flows)
    out2 = ops.warp(_t(imgs[:1], cuda), flow, 'cv2_cubic', -1.0).cpu().numpy()
    staged2, fallback2 = ops.warp_tile_stats(reset=True)
    assert staged2 >= 0.95 * ntiles
    assert np.array_equal(out2[1], wo.warp_frame_raft(imgs[0], fl[1]))
    # and the noisy flow of the test above really is the fallback path
    noisy = torch.randn((1, H, W, 2), device=cuda) * 12
    ops.warp(_t(imgs[:1], cuda), noisy, 'cv2_cubic', 1.0)
    s3, f3 = ops.warp_tile_stats(reset=True)
    print(f'noisy 12 px flow: {f3} of {s3 + f3} tiles fall back')


def test_cubic_u8_batched_and_shared_source(cuda):
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(8)
    B, H, W = 3, 40, 36
    imgs = rs.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    flows = (5.0 * rs.standard_normal((B, H, W, 2))).astype(np.float32)
    out = ops.warp(_t(imgs, cuda), _t(flows, cuda)).cpu().numpy()
    for b in range(B):
        assert np.array_equal(out[b], wo.warp_frame_pdcnet(imgs[b], flows[b]))
    shared = ops.warp(_t(imgs[:1], cuda), _t(flows, cuda)).cpu().numpy()   # one key frame, B flows
    for b in range(B):
        assert np.array_equal(shared[b], wo.warp_frame_pdcnet(imgs[0], flows[b]))
    # different source and destination sizes
    big = rs.randint(0, 256, (64, 80, 3)).astype(np.uint8)
    fl = (6.0 * rs.standard_normal((20, 24, 2))).astype(np.float32) + 10
    mx, my = wo.maps_pdcnet(fl)
    assert np.array_equal(ops.warp(_t(big, cuda), _t(fl, cuda)).cpu().numpy(), wo.remap_cubic(big, mx, my))


@pytest.mark.parametrize('C', [1, 2, 4])
def test_cubic_generic_channel_counts(cuda, C):
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(C)
    img = rs.randint(0, 256, (33, 47, C)).astype(np.uint8)
    flow = (7.0 * rs.standard_normal((33, 47, 2))).astype(np.float32)
    out = ops.warp(_t(img, cuda), _t(flow, cuda)).cpu().numpy()
    assert np.array_equal(out, wo.warp_frame_pdcnet(img, flow))
    imgf = rs.standard_normal((33, 47, C)).astype(np.float32)
    outf = ops.warp(_t(imgf, cuda), _t(flow, cuda)).cpu().numpy()
    np.testing.assert_allclose(outf, wo.warp_frame_pdcnet(imgf, flow), rtol=0, atol=1e-5)


def test_unaligned_source_pointer_takes_the_tap_loop(cuda):
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(2)
    raw = torch.from_numpy(rs.randint(0, 256, (1 + 24 * 28 * 3,)).astype(np.uint8)).to(cuda)
    img = raw[1:].view(24, 28, 3)      # data_ptr is odd
    flow = (3.0 * rs.standard_normal((24, 28, 2))).astype(np.float32)
    assert img.data_ptr() % 4 != 0
    out = ops.warp(img, _t(flow, cuda)).cpu().numpy()
    assert np.array_equal(out, wo.warp_frame_pdcnet(img.cpu().numpy(), flow))


def test_bilinear_vs_oracle_and_grid_sample(cuda):
    from sd_animation_optical_flow_b200 import ops
    img = gi.texture(96, 80, 5).astype(np.float32)
    rs = np.random.RandomState(3)
    flow = (4.0 * rs.standard_normal((96, 80, 2))).astype(np.float32)
    flow[0, 0] = [np.nan, 1e20]
    d_img, d_flow = _t(img, cuda), _t(flow, cuda)
    out = ops.warp(d_img, d_flow, 'bilinear').cpu().numpy()
    ref = wo.warp_bilinear(img, np.nan_to_num(flow, nan=1e9, posinf=1e9))
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-4)        # tolerance: 1e-4 abs on [0,255]
    H, W = flow.shape[:2]
    xs = torch.arange(W, device=cuda).float()[None, :] + d_flow[..., 0]
    ys = torch.arange(H, device=cuda).float()[:, None] + d_flow[..., 1]
    grid = torch.stack([2 * xs / (W - 1) - 1, 2 * ys / (H - 1) - 1], -1)[None]
    gs = torch.nn.functional.grid_sample(d_img.permute(2, 0, 1)[None], grid, mode='bilinear', padding_mode='zeros',
                                         align_corners=True)[0].permute(1, 2, 0).cpu().numpy()
    ok = np.isfinite(gs)
    ok[0, 0] = False
    np.testing.assert_allclose(out[ok], gs[ok], rtol=0, atol=5e-3)  # grid_sample's coordinate round trip
    u8 = ops.warp(_t(gi.texture(96, 80, 5), cuda), d_flow, 'bilinear').cpu().numpy()
    assert np.array_equal(u8, wo.warp_bilinear(gi.texture(96, 80, 5), np.nan_to_num(flow, nan=1e9, posinf=1e9)))


def test_zero_flow_is_identity_at_config5_size(cuda):
    from sd_animation_optical_flow_b200 import ops
    img = torch.randint(0, 256, (2, 720, 1280, 3), dtype=torch.uint8, device=cuda)
    flow = torch.zeros((2, 720, 1280, 2), device=cuda)
    assert torch.equal(ops.warp(img, flow), img)
    assert torch.equal(ops.warp(img, flow, sign=-1.0), img)
    assert torch.equal(ops.warp(img, flow, 'bilinear'), img)
    # integer translation = shifted copy with a zero border
    flow[..., 0] = 5.0
    flow[..., 1] = -3.0
    out = ops.warp(img, flow)
    assert torch.equal(out[:, 3:, :-5], img[:, :-3, 5:])
    assert int(out[:, :3].max()) == 0 and int(out[:, :, -5:].max()) == 0


def test_warp_frame_latent(cuda):
    """W3 with no host arithmetic: both cv2.resize(INTER_CUBIC) calls of pdcnet_of.py:19-32 run in csrc/resize.cu.  Against
    the oracle (which calls cv2.resize itself): 2e-5 abs on unit-variance latents, as for the float warp."""
    from sd_animation_optical_flow_b200 import pdcnet_of
    rs = np.random.RandomState(4)
    for (lh, lw, C) in ((12, 16, 4), (96, 64, 4), (10, 14, 3)):
        lat = torch.from_numpy(rs.standard_normal((1, C, lh, lw)).astype(np.float32))
        flow = (3.0 * rs.standard_normal((8 * lh, 8 * lw, 2))).astype(np.float32)
        out = pdcnet_of.warp_frame_latent(lat, flow)
        ref = wo.warp_frame_latent(lat[0].numpy(), flow)
        assert out.shape == (1, C, lh, lw) and out.device.type == 'cpu' and out.dtype == torch.float32
        np.testing.assert_allclose(out[0].numpy(), ref, rtol=0, atol=2e-5)
        assert torch.equal(pdcnet_of.warp_frame_latent(lat.to(cuda), flow), out)       # CUDA latents are taken as they are


@pytest.mark.parametrize('case', [(12, 16, 4, 96, 128), (96, 128, 4, 12, 16), (96, 64, 4, 768, 512), (768, 512, 4, 96, 64),
                                  (17, 23, 3, 50, 41), (50, 41, 1, 17, 23), (33, 20, 6, 100, 7), (5, 7, 2, 5, 7)])
def test_resize_cubic_f32_vs_cv2_and_oracle(cuda, case):
    """csrc/resize.cu against the NumPy restatement (same operation order: 1e-6 of the range) and against cv2.resize itself
    (3e-7 at the integer ratios the reference uses, 3e-6 at arbitrary ratios, see tests/test_oracle_warp.py)."""
    import cv2
    from sd_animation_optical_flow_b200 import ops
    hs, ws, c, hd, wd = case
    rs = np.random.RandomState(hs + wd)
    img = (3 * rs.standard_normal((2, hs, ws, c))).astype(np.float32)
    out = ops.resize_cubic(_t(img, cuda), hd, wd).cpu().numpy()
    rng = float(np.abs(img).max())
    integer_ratio = (hd % hs == 0 and wd % ws == 0) or (hs % hd == 0 and ws % wd == 0)
    for b in range(2):
        ora = wo.resize_cubic_f32(img[b], (wd, hd))
        ref = cv2.resize(img[b], (wd, hd), interpolation=cv2.INTER_CUBIC).reshape(hd, wd, c)
        assert np.abs(out[b] - ora).max() <= 1e-6 * rng
        assert np.abs(out[b] - ref).max() <= (3e-7 if integer_ratio else 3e-6) * rng
    one = ops.resize_cubic(_t(img[0, :, :, 0], cuda), hd, wd)
    assert one.shape == (hd, wd) and torch.equal(one, _t(out[0, :, :, 0], cuda))
