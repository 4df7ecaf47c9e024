"""GPU parity of confidence / mask / composite kernels: integer results are
bit-exact vs the reference's outputs (golden) and vs the oracle on larger seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import mask_oracle as mo
from oracle import warp_oracle as wo
from tests import golden_inputs as gi

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_generate_mask_golden(cuda, golden):
    from sd_animation_optical_flow_b200 import ofgen
    conf, logc, *_ = gi.mask_inputs()
    for thres in (0.5, 0.95):
        lc = logc.copy()
        m, lc_out = ofgen.generate_mask(conf, lc, thres)
        assert np.array_equal(m, golden['mask'][f'mask_{thres}'])
        assert lc_out is lc and np.array_equal(lc, golden['mask'][f'logc_{thres}'])   # in-place like the reference


def test_dilate_expand_golden(cuda, golden):
    from sd_animation_optical_flow_b200 import ofgen, ops
    conf, _, img, *_ = gi.mask_inputs()
    m = golden['mask']['mask_0.5']
    low = (conf < 0.2).astype(np.uint8) * 255
    assert np.array_equal(ops.dilate_ellipse(_t(low, cuda)[None], 15)[0].cpu().numpy(), golden['mask']['dilate15'])
    assert np.array_equal(ofgen.invert_and_dilate(m), golden['mask']['invert_dilate'])
    assert np.array_equal(ofgen.expand_mask(m, img), golden['mask']['expand'])


def test_mix_merge_travel_golden(cuda, golden):
    from sd_animation_optical_flow_b200 import ofgen, ops
    conf, _, _, raw, warped = gi.mask_inputs()
    m = golden['mask']['mask_0.5']
    for ppw in (1.0, 0.3):
        assert np.array_equal(ofgen.mix_propagated_ai_frame(raw, warped, m, ppw), golden['mask'][f'mix_{ppw}'])
    assert ofgen.mix_propagated_ai_frame(raw, warped, m, 0.0) is raw
    assert np.array_equal(ofgen.merge_images(raw, warped, m), golden['mask']['merge'])
    flow = gi.warp_inputs('small_u8')[1][: conf.shape[0], : conf.shape[1]].copy()
    v = ops.travel_distance(_t(flow, cuda)[None], _t(conf, cuda)[None], 0.9)[0].cpu().numpy()
    assert np.array_equal(v, golden['mask']['travel'])


@pytest.mark.parametrize('ksize', [1, 3, 7, 15, 31])
@pytest.mark.parametrize('hw', [(768, 512), (67, 129)])
def test_dilate_any_u8_vs_oracle(cuda, ksize, hw):
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(ksize)
    g = rs.randint(0, 256, (2, *hw)).astype(np.uint8)
    g[rs.uniform(size=g.shape) < 0.9] = 0
    out = ops.dilate_ellipse(_t(g, cuda), ksize).cpu().numpy()
    inv = ops.dilate_ellipse(_t(g, cuda), ksize, invert=True).cpu().numpy()
    for b in range(2):
        assert np.array_equal(out[b], mo.dilate_ellipse(g[b], ksize))
        assert np.array_equal(inv[b], mo.dilate_ellipse(255 - g[b], ksize))


def test_softmax_confidence(cuda):
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(5)
    wm = (2 * rs.standard_normal((3, 2, 37, 41))).astype(np.float32)
    conf, logc = ops.confidence_softmax(_t(wm, cuda))
    rc, rl = mo.confidence_from_weight_map(wm)
    np.testing.assert_allclose(conf.cpu().numpy(), rc, rtol=2e-6, atol=1e-7)     # tolerance: 2 ulp of fp32 exp/div
    np.testing.assert_allclose(logc.cpu().numpy(), rl, rtol=2e-6, atol=1e-6)
    t = torch.from_numpy(wm)
    np.testing.assert_allclose(conf.cpu().numpy(), t.softmax(1)[:, 0].numpy(), rtol=3e-6, atol=1e-7)


def test_greedy_composite_golden(cuda, golden):
    from sd_animation_optical_flow_b200 import ofgen
    fm, frames, thres = gi.greedy_inputs()
    ret, mask, order = ofgen.composite_references(fm, list(frames), thres)
    assert order == golden['greedy']['order'].tolist()
    assert np.array_equal(ret, golden['greedy']['ret'])
    assert np.array_equal(mask, golden['greedy']['mask'])
    assert np.array_equal(fm[..., 2], golden['greedy']['conf_after'])      # flow_mat updated in place like the reference


def test_greedy_composite_large_vs_oracle(cuda):
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(11)
    n, H, W = 6, 96, 128
    fm = np.zeros((n, 1, H, W, 3), np.float32)
    fm[..., :2] = (4 * rs.standard_normal((n, 1, H, W, 2))).astype(np.float32)
    fm[..., 2] = rs.uniform(0, 1, (n, 1, H, W)).astype(np.float32)
    fm[2, 0, :, :, 2] = 0      # a reference that never wins
    frames = rs.randint(0, 256, (n, H, W, 3)).astype(np.uint8)
    ref_ret, ref_mask, ref_order = mo.greedy_composite(fm, list(frames), 0.6)
    d_fm = _t(fm.reshape(n, H, W, 3), cuda).clone()
    ret, mask, order = ops.greedy_composite(d_fm, _t(frames, cuda), 0.6)
    assert order.cpu().tolist() == ref_order
    assert np.array_equal(ret.cpu().numpy(), ref_ret) and np.array_equal(mask.cpu().numpy(), ref_mask)


def test_confidence_sums_and_keyframe_pick(cuda):
    from sd_animation_optical_flow_b200 import ofgen, ops
    rs = np.random.RandomState(12)
    fm = rs.uniform(0, 1, (5, 5, 64, 48, 3)).astype(np.float32)
    sums = ops.confidence_sums(_t(fm, cuda)).cpu().numpy()
    ref = mo.keyframe_scores(fm)
    np.testing.assert_allclose(sums, ref, rtol=1e-12)
    assert ofgen.keyframe_conv_pick(fm) == int(np.argmax(ref))


def test_fused_warp_mask_composite_vs_oracle(cuda):
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(13)
    B, H, W = 2, 150, 131
    src = rs.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    base = rs.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    flow = (5 * rs.standard_normal((B, H, W, 2))).astype(np.float32)
    wm = np.stack([gi._blur(rs.standard_normal((H, W)) * 6, 2.0) * 6 for _ in range(2 * B)]).reshape(B, 2, H, W).astype(np.float32)
    thres = 0.5
    out, mask = ops.warp_mask_composite(_t(src, cuda), _t(base, cuda), _t(flow, cuda), _t(wm, cuda), thres, 7)
    out, mask = out.cpu().numpy(), mask.cpu().numpy()
    conf, _ = mo.confidence_from_weight_map(wm)
    for b in range(B):
        # pixels whose confidence is within 1e-5 of the threshold may legitimately flip (fp32 exp ulp);
        # compare everywhere their 7x7 neighbourhood cannot reach
        unsure = (np.abs(conf[b] - thres) < 1e-5).astype(np.uint8) * 255
        safe = mo.dilate_ellipse(unsure, 7) == 0
        m_ref, _ = mo.generate_mask(conf[b], np.zeros_like(conf[b]), thres, 7)
        warped = wo.warp_frame_pdcnet(src[b], flow[b])
        ref = mo.mix_propagated_ai_frame(base[b], warped, m_ref, 1.0)
        assert safe.mean() > 0.99
        assert np.array_equal(mask[b][safe], m_ref[safe])
        assert np.array_equal(out[b][safe], ref[safe])
    # shared key frame
    out1, _ = ops.warp_mask_composite(_t(src[:1], cuda), _t(base, cuda), _t(flow, cuda), _t(wm, cuda), thres, 7)
    assert np.array_equal(out1[0].cpu().numpy(), out[0])
