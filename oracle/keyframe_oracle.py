"""CPU oracle for the key-frame detector: the reference's own expressions
(ofgen_pixel_inpaint.py:127-176, 272-313), which are OpenCV / NumPy calls -- kept literal here, so the oracle IS the
reference's arithmetic running on the real cv2.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import math

import cv2
import numpy as np


def mean_pixel_distance(left: np.ndarray, right: np.ndarray) -> float:
    """ofgen_pixel_inpaint.py:132-139."""
    assert len(left.shape) == 2 and len(right.shape) == 2
    assert left.shape == right.shape
    num_pixels = float(left.shape[0] * left.shape[1])
    return np.sum(np.abs(left.astype(np.int32) - right.astype(np.int32))) / num_pixels


def estimated_kernel_size(frame_width: int, frame_height: int) -> int:
    """ofgen_pixel_inpaint.py:142-147."""
    size = 4 + round(math.sqrt(frame_width * frame_height) / 192)
    if size % 2 == 0:
        size += 1
    return size


def thresholds(lum: np.ndarray):
    """ofgen_pixel_inpaint.py:158-161."""
    sigma = 1.0 / 3.0
    median = np.median(lum)
    return int(max(0, (1.0 - sigma) * median)), int(min(255, (1.0 + sigma) * median))


def detect_edges(frame: np.ndarray) -> np.ndarray:
    """ofgen_pixel_inpaint.py:150-176 (the kernel is rebuilt per call instead of cached in a module global)."""
    hue, sat, lum = cv2.split(cv2.cvtColor(frame, cv2.COLOR_BGR2HSV))
    k = estimated_kernel_size(lum.shape[1], lum.shape[0])
    low, high = thresholds(lum)
    edges = cv2.Canny(lum, low, high)
    return cv2.dilate(edges, np.ones((k, k), np.uint8))


def key_frame_flags(frames, fps: float = 30.0, th: float = 8.5, keep_every: int = 1):
    """The decision sequence of frame_generator (ofgen_pixel_inpaint.py:272-313) for already-resized frames."""
    min_gap = int(10 * fps / 30)  # noqa: F841  (computed but unused by the reference as well)
    max_gap = int(300 * fps / 30)
    gap, key_edges, flags = 0, None, []
    for frame in frames:
        gap += keep_every
        edges = detect_edges(frame)
        if key_edges is None:
            key_edges = edges
            flags.append(True)
            continue
        delta = mean_pixel_distance(edges, key_edges)
        if th * (max_gap - gap) / max_gap < delta:
            key_edges = edges
            gap = 0
            flags.append(True)
        else:
            flags.append(False)
    return flags
