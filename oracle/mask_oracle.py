"""Oracle for confidence post-processing, mask generation and composites
(SURVEY §8a rows M1-M6, A1 identity pairs, A2).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  NumPy restatements of the
reference's numpy/cv2 code; pinned against cv2 (dilate, Laplacian, cvtColor,
getStructuringElement) and against the reference functions' literal numpy
expressions in tests/test_oracle_mask.py.
"""
from __future__ import annotations

import numpy as np

from . import warp_oracle


# ----------------------------------------------------------------------------- M1
def confidence_from_weight_map(weight_map: np.ndarray):
    """PDCNetPlus.calc post-processing (pdcnet_of.py:72-74).
    weight_map [B,K,H,W] logits -> (confidence, log_confidence) = component 0 of
    softmax / log_softmax over K, each [B,H,W] fp32."""
    w = weight_map.astype(np.float32)
    m = w.max(axis=1, keepdims=True)
    e = np.exp(w - m).astype(np.float32)
    s = e.sum(axis=1, keepdims=True, dtype=np.float32)
    conf = (e / s)[:, 0].astype(np.float32)
    logc = ((w - m) - np.log(s))[:, 0].astype(np.float32)
    return conf, logc


# ----------------------------------------------------------------------------- M2
def travel_distance(flow: np.ndarray, confidence: np.ndarray, conf_thres: float = 0.9) -> np.ndarray:
    """of_calc (ofgen_pixel_inpaint.py:105-118): |displacement| recomputed
    through the float32 map round trip, zeroed where confidence < 0.9."""
    h, w = flow.shape[:2]
    mx, my = warp_oracle.maps_pdcnet(flow)
    mx = mx.copy()
    my = my.copy()
    mx -= np.arange(w)
    my -= np.arange(h)[:, np.newaxis]
    v = np.sqrt(mx * mx + my * my)
    v[confidence < conf_thres] = 0
    return v


# ----------------------------------------------------------------------------- dilate
def ellipse_half_widths(ksize: int):
    """Row half-widths of cv2.getStructuringElement(MORPH_ELLIPSE, (k,k)):
    dx = round(c*sqrt((r^2-dy^2)/r^2)), r=c=k//2 (rows with |dy|<=r)."""
    r = ksize // 2
    c = ksize // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    hw = []
    for i in range(ksize):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))
            hw.append(dx)
        else:
            hw.append(-1)
    return hw


def dilate_ellipse(mask: np.ndarray, ksize: int = 7) -> np.ndarray:
    """cv2.dilate(mask, getStructuringElement(MORPH_ELLIPSE,(k,k))) with the
    default border (outside pixels never win the max)."""
    H, W = mask.shape
    r = ksize // 2
    hw = ellipse_half_widths(ksize)
    out = np.zeros_like(mask)
    pad = np.zeros((H + 2 * r, W + 2 * r), mask.dtype)
    pad[r:r + H, r:r + W] = mask
    for i, half in enumerate(hw):
        if half < 0:
            continue
        for dx in range(-half, half + 1):
            out = np.maximum(out, pad[i:i + H, r + dx:r + dx + W])
    return out


# ----------------------------------------------------------------------------- M3
def generate_mask(confidence: np.ndarray, log_confidence: np.ndarray, thres: float = 0.8, ksize: int = 7):
    """generate_mask (ofgen_pixel_inpaint.py:262-267, ofgen_keyframe_inpaint.py:317-322).
    Returns (dilated mask u8, log_confidence with masked pixels reset to 0)."""
    mask = np.zeros(confidence.shape, np.uint8)
    low = confidence < thres
    mask[low] = 255
    logc = log_confidence.copy()
    logc[low] = 0
    return dilate_ellipse(mask, ksize), logc


def confidence_to_mask(confidence, flow, dist, pixel_travel_dist, travel_thres, conf_thres=0.9, ksize=15):
    """confidence_to_mask (ofgen_pixel_inpaint.py:218-227; unused by the live
    scripts).  Returns (mask u8, updated pixel_travel_dist fp32)."""
    mask = np.zeros(confidence.shape, np.uint8)
    mask[confidence < conf_thres] = 255
    ptd = warp_oracle.warp_frame_pdcnet(pixel_travel_dist.astype(np.float32), flow) + dist
    ptd[confidence < conf_thres] = 0
    mask[ptd > travel_thres] = 255
    ptd[ptd > travel_thres] = 0
    return dilate_ellipse(mask, ksize), ptd


# ----------------------------------------------------------------------------- M4
def mix_propagated_ai_frame(raw: np.ndarray, warped: np.ndarray, mask: np.ndarray, ppw: float = 1.0) -> np.ndarray:
    """mix_propagated_ai_frame (ofgen_pixel_inpaint.py:251-260)."""
    if ppw < 0.001:
        return raw
    weights = np.zeros(raw.shape[:2], np.float32)
    weights[mask <= 127] = ppw
    weights[mask > 127] = 1 - ppw
    weights = weights[:, :, None]
    out = raw.astype(np.float32) * (1 - weights) + warped.astype(np.float32) * weights
    return np.clip(out, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------- M5
def merge_images(base: np.ndarray, second: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """merge_images(method='naive') (ofgen_keyframe_inpaint.py:676-681):
    (mask/255).astype(u8) is 1 only where mask == 255 -> a select."""
    m = (mask == 255)[:, :, None]
    return np.where(m, second, base).astype(base.dtype)


def greedy_composite(flow_mat: np.ndarray, ai_frames, thres: float):
    """Greedy multi-reference composite
    (ofgen_keyframe_inpaint.py:995-1024; same loop at :741-770).
    flow_mat [n,1,H,W,3] fp32 (flow x, flow y, confidence); ai_frames: n u8
    [H,W,3].  Returns (ret u8 [H,W,3], mask u8 [H,W], order of chosen refs)."""
    fm = flow_mat.astype(np.float32).copy()
    n = fm.shape[0]
    fm[..., 2] = (fm[..., 2] > thres).astype(np.float32)
    H, W = fm.shape[2:4]
    mask = np.zeros((H, W), np.uint8)
    ret = None
    order = []
    for _ in range(n):
        sums = fm[..., 2].reshape(n, -1).sum(axis=1, dtype=np.float64)
        ref = int(np.argmax(sums))
        order.append(ref)
        warped = warp_oracle.warp_frame_pdcnet(ai_frames[ref], fm[ref, 0, :, :, 0:2])
        last = fm[ref, 0, :, :, 2].copy()
        cur_mask = (last * 255).astype(np.uint8)
        mask = mask | cur_mask
        ret = warped.copy() if ret is None else merge_images(ret, warped, cur_mask)
        fm[:, 0, :, :, 2] -= last[None]
        fm[:, 0, :, :, 2] = np.clip(fm[:, 0, :, :, 2], 0, 1)
    return ret, mask, order


# ----------------------------------------------------------------------------- M6
def laplacian_abs_u8(img: np.ndarray) -> np.ndarray:
    """np.absolute(cv2.Laplacian(img, CV_64F)).astype(np.uint8): 3x3
    [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101, |.| then wrap mod 256."""
    x = img.astype(np.int64)
    p = np.pad(x, ((1, 1), (1, 1), (0, 0)), mode='reflect')
    lap = p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:] - 4 * p[1:-1, 1:-1]
    return (np.abs(lap) & 0xFF).astype(np.uint8)


def rgb2gray_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, COLOR_RGB2GRAY) for u8: 15-bit fixed point."""
    x = img.astype(np.int64)
    return ((x[..., 0] * 9798 + x[..., 1] * 19235 + x[..., 2] * 3735 + 16384) >> 15).astype(np.uint8)


def expand_mask(mask: np.ndarray, ori_image: np.ndarray, ksize: int = 7) -> np.ndarray:
    """expand_mask (ofgen_keyframe_inpaint.py:968-973)."""
    lap = (rgb2gray_u8(laplacian_abs_u8(ori_image)) > 20).astype(np.uint8) * 255
    lap = dilate_ellipse(lap, ksize)
    return mask | lap


def invert_dilate(mask: np.ndarray, ksize: int = 7) -> np.ndarray:
    """mask2 = dilate(255 - mask) (ofgen_keyframe_inpaint.py:772-774)."""
    return dilate_ellipse((255 - mask).astype(np.uint8), ksize)


# ----------------------------------------------------------------------------- A2
def keyframe_scores(flow_mat: np.ndarray) -> np.ndarray:
    """KeyframeConv score (ofgen_keyframe_inpaint.py:664-668): sum_{t,h,w} conf[s,t]."""
    n = flow_mat.shape[0]
    return flow_mat[..., 2].reshape(n, -1).sum(axis=1, dtype=np.float64)
