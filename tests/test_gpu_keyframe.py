"""GPU parity of the key-frame detector through the C ABI: bit-exact against the reference's own
OpenCV expressions (oracle/keyframe_oracle.py = ofgen_pixel_inpaint.py:127-176 verbatim)."""
import cv2
import numpy as np
import pytest
import torch

from oracle import keyframe_oracle as ko
from tests import golden_inputs as gi

pytestmark = pytest.mark.gpu


def _frames(H, W, n, seed):
    """A blurred texture with flat shapes on it (so Canny finds sparse, strong edges); consecutive frames drift by a
    pixel every 4 frames, every 4th frame jumps to a random crop (a scene change)."""
    rs = np.random.RandomState(seed)
    base = cv2.GaussianBlur(gi.texture(H + 64, W + 64, seed), (0, 0), 3.0)
    for _ in range(25):
        x, y = rs.randint(0, W + 40), rs.randint(0, H + 40)
        w, h = rs.randint(8, 60, 2)
        col = tuple(int(c) for c in rs.randint(0, 256, 3))
        if rs.rand() < 0.5:
            cv2.rectangle(base, (x, y), (x + w, y + h), col, -1)
        else:
            cv2.circle(base, (x, y), int(w) // 2, col, -1)
    out = []
    for i in range(n):
        dy, dx = (rs.randint(0, 64, 2) if i % 4 == 3 else (i // 4, (i // 4) % 64))
        out.append(np.ascontiguousarray(base[dy:dy + H, dx:dx + W, ::-1]))
    return out


@pytest.mark.parametrize('hw', [(768, 512), (720, 1280), (61, 45), (33, 200)])
def test_detect_edges_bit_exact_vs_opencv(cuda, hw):
    from sd_animation_optical_flow_b200 import ofgen, ops
    H, W = hw
    for seed, frame in enumerate(_frames(H, W, 3, H)):
        ref = ko.detect_edges(frame)
        got = ofgen.detect_edges(frame)
        assert got.shape == ref.shape and got.dtype == np.uint8
        assert np.array_equal(got, ref), f'{(got != ref).sum()} of {ref.size} pixels differ'
        # the undilated Canny map with explicit thresholds
        lum = frame.max(axis=2)
        can = ops.detect_edges(torch.from_numpy(frame).to(cuda), 1, 40, 120).cpu().numpy()
        assert np.array_equal(can, cv2.Canny(lum, 40, 120))


def test_noise_image_and_swapped_thresholds(cuda):
    """Dense edges (long hysteresis chains) and low > high (OpenCV swaps them)."""
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(1)
    frame = rs.randint(0, 256, (200, 300, 3)).astype(np.uint8)
    frame = cv2.GaussianBlur(frame, (0, 0), 1.0)
    lum = frame.max(axis=2)
    for low, high in ((20, 60), (90, 30), (0, 255), (5, 5)):
        got = ops.detect_edges(torch.from_numpy(frame).to(cuda), 1, low, high).cpu().numpy()
        assert np.array_equal(got, cv2.Canny(lum, low, high)), (low, high)


def test_mean_pixel_distance_and_selector(cuda):
    from sd_animation_optical_flow_b200 import ofgen
    frames = _frames(192, 160, 12, 7)
    e0, e1 = ko.detect_edges(frames[0]), ko.detect_edges(frames[3])
    assert ofgen.mean_pixel_distance(e0, e1) == ko.mean_pixel_distance(e0, e1)
    sel = ofgen.KeyFrameSelector(fps=30.0, th=8.5)
    flags = [sel.push(f, frames_advanced=3) for f in frames]
    assert flags == ko.key_frame_flags(frames, fps=30.0, th=8.5, keep_every=3)
    assert flags[0] and any(flags[1:]) and not all(flags)
    assert ofgen.estimated_kernel_size(512, 768) == ko.estimated_kernel_size(512, 768) == 7
