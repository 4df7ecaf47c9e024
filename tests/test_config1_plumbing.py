"""BASELINE.json configs[0]: "2-frame 256x256 cv2 Farneback CPU flow + torch.grid_sample warp (plumbing, no GPU)".

The CPU reference arm of bench.py (`--impl reference`, `cpu_baseline`) on the synthetic pair SURVEY §8(d) defines: a
Gaussian-blurred noise texture and the same texture shifted by (+4, -3) px.  Checks the plumbing (shapes, dtypes, the
BGR/gray conversions) and that the recipe recovers the shift: Farneback's flow of frame1 -> frame2 points from a pixel of
frame 1 to where its content sits in frame 2, i.e. (+4, -3); warping frame 2 back with it reproduces frame 1.
"""
import numpy as np

from oracle import farneback_baseline as fb
from tests import golden_inputs as gi


def test_config1_farneback_grid_sample_recovers_the_shift():
    f1, f2 = gi.shifted_pair(256, 256, 0)              # RGB uint8; frame 2 = frame 1 shifted by (+4, -3)
    a, b = f1[:, :, ::-1].copy(), f2[:, :, ::-1].copy()
    flow, warped = fb.flow_and_warp(a, b, b)
    assert flow.shape == (256, 256, 2) and flow.dtype == np.float32
    assert warped.shape == (256, 256, 3) and warped.dtype == np.float32
    inner = (slice(24, -24), slice(24, -24))
    med = np.median(flow[inner].reshape(-1, 2), axis=0)
    assert abs(med[0] - 4.0) <= 0.25 and abs(med[1] + 3.0) <= 0.25, med
    err = np.abs(flow[inner] - np.array([4.0, -3.0], np.float32)).mean()
    assert err <= 0.5, err
    # frame 2 sampled at x + flow(x) reproduces frame 1 (bilinear, away from the border)
    diff = np.abs(warped[inner] - a[inner].astype(np.float32)).mean()
    assert diff <= 4.0, diff
    # the timed form used by bench.py returns a rate and per-pair times
    rate, n, times = fb.time_pairs(a, b, b, budget_s=0.2, min_pairs=2, max_pairs=5)
    assert rate > 0 and n == len(times) >= 2
    info = fb.host_info()
    assert info['cores'] >= 1 and info['cv2_threads'] >= 1
