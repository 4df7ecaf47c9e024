"""Pins oracle/mask_oracle.py against the reference's outputs (tests/golden/mask.npz and greedy.npz,
generated with the literal numpy/cv2 expressions of ofgen_pixel_inpaint.py / ofgen_keyframe_inpaint.py)."""
import numpy as np
import pytest

from oracle import mask_oracle as mo
from tests import golden_inputs as gi


def test_generate_mask(golden):
    conf, logc, *_ = gi.mask_inputs()
    for thres in (0.5, 0.95):
        m, lc = mo.generate_mask(conf, logc, thres)
        assert np.array_equal(m, golden['mask'][f'mask_{thres}'])
        assert np.array_equal(lc, golden['mask'][f'logc_{thres}'])


def test_dilations(golden):
    conf, *_ = gi.mask_inputs()
    assert np.array_equal(mo.dilate_ellipse((conf < 0.2).astype(np.uint8) * 255, 15), golden['mask']['dilate15'])
    assert np.array_equal(mo.invert_dilate(golden['mask']['mask_0.5'], 7), golden['mask']['invert_dilate'])


def test_ellipse_rows_vs_cv2():
    cv2 = pytest.importorskip('cv2')
    for k in (1, 3, 5, 7, 9, 15, 21, 31):
        kern = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
        hw = mo.ellipse_half_widths(k)
        mine = np.zeros((k, k), np.uint8)
        for i, h in enumerate(hw):
            mine[i, k // 2 - h:k // 2 + h + 1] = 1
        assert np.array_equal(mine, kern), k


def test_expand_mask(golden):
    _, _, img, *_ = gi.mask_inputs()
    assert np.array_equal(mo.expand_mask(golden['mask']['mask_0.5'], img), golden['mask']['expand'])


def test_mix_and_merge(golden):
    _, _, _, raw, warped = gi.mask_inputs()
    m = golden['mask']['mask_0.5']
    for ppw in (1.0, 0.3):
        assert np.array_equal(mo.mix_propagated_ai_frame(raw, warped, m, ppw), golden['mask'][f'mix_{ppw}'])
    assert mo.mix_propagated_ai_frame(raw, warped, m, 0.0) is raw
    assert np.array_equal(mo.merge_images(raw, warped, m), golden['mask']['merge'])


def test_travel_distance(golden):
    conf, *_ = gi.mask_inputs()
    flow = gi.warp_inputs('small_u8')[1][: conf.shape[0], : conf.shape[1]].copy()
    assert np.array_equal(mo.travel_distance(flow, conf, 0.9), golden['mask']['travel'])


def test_greedy_composite(golden):
    fm, frames, thres = gi.greedy_inputs()
    ret, mask, order = mo.greedy_composite(fm, list(frames), thres)
    assert order == golden['greedy']['order'].tolist()
    assert np.array_equal(ret, golden['greedy']['ret'])
    assert np.array_equal(mask, golden['greedy']['mask'])


def test_softmax_confidence_vs_torch():
    torch = pytest.importorskip('torch')
    rs = np.random.RandomState(5)
    wm = (2 * rs.standard_normal((2, 2, 9, 11))).astype(np.float32)
    conf, logc = mo.confidence_from_weight_map(wm)
    t = torch.from_numpy(wm)
    np.testing.assert_allclose(conf, t.softmax(dim=1)[:, 0].numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(logc, t.log_softmax(dim=1)[:, 0].numpy(), rtol=2e-6, atol=1e-6)


def test_keyframe_scores():
    fm, _, _ = gi.greedy_inputs()
    s = mo.keyframe_scores(fm)
    assert s.shape == (fm.shape[0],)
    np.testing.assert_allclose(s, fm[..., 2].reshape(fm.shape[0], -1).astype(np.float64).sum(1))
