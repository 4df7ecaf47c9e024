"""Oracle for the backward warp (SURVEY §8a rows W1, W2, W3 and the bilinear mode).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference warps every image with
    cv2.remap(frame, map_x, map_y, INTER_CUBIC, BORDER_CONSTANT)     pdcnet_of.py:34-42
    cv2.remap(frame, map_xy, None, INTER_CUBIC)                       ofgen.py:37-43
so the "algorithm" is OpenCV's fixed-point bicubic remap.  It is restated here
from its published behaviour (imgproc: initInterTab2D / remapBicubic, OpenCV
4.x) and pinned bit-exactly against cv2.remap in tests/test_oracle_warp.py.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS          # 32 sub-pixel positions per axis
INTER_REMAP_COEF_BITS = 15
INTER_REMAP_COEF_SCALE = 1 << INTER_REMAP_COEF_BITS


def cubic_coeffs_1d() -> np.ndarray:
    """float32 [32,4] bicubic (a=-0.75) taps for fractions i/32, evaluated in
    fp32 in OpenCV's operation order (c3 = 1 - c0 - c1 - c2)."""
    f = np.float32
    A = f(-0.75)
    tab = np.zeros((INTER_TAB_SIZE, 4), np.float32)
    scale = f(1.0) / f(INTER_TAB_SIZE)
    for i in range(INTER_TAB_SIZE):
        x = f(i) * scale
        xp1 = f(x + f(1))
        c0 = f(f(f(f(f(A * xp1) - f(f(5) * A)) * xp1) + f(f(8) * A)) * xp1) - f(f(4) * A)
        c0 = f(c0)
        c1 = f(f(f(f(f(f(A + f(2)) * x) - f(A + f(3))) * x) * x) + f(1))
        omx = f(f(1) - x)
        c2 = f(f(f(f(f(f(A + f(2)) * omx) - f(A + f(3))) * omx) * omx) + f(1))
        c3 = f(f(f(f(1) - c0) - c1) - c2)
        tab[i] = (c0, c1, c2, c3)
    return tab


def cubic_table_i16() -> np.ndarray:
    """int16 [1024,16]: entry (fy*32+fx) holds the 4x4 weights w[ky*4+kx],
    scaled by 2^15, rounded half-to-even, and fixed up so they sum to 2^15
    exactly (the correction goes to the max (sum too small) or min (sum too
    large) of the four centre-right/bottom taps ky,kx in {2,3})."""
    t1 = cubic_coeffs_1d()
    out = np.zeros((INTER_TAB_SIZE * INTER_TAB_SIZE, 16), np.int16)
    for fy in range(INTER_TAB_SIZE):
        for fx in range(INTER_TAB_SIZE):
            v = (t1[fy][:, None] * t1[fx][None, :]).astype(np.float32)           # fp32 product
            q = np.rint(v * np.float32(INTER_REMAP_COEF_SCALE)).astype(np.int64)  # cvRound
            q = np.clip(q, -32768, 32767).astype(np.int32)
            isum = int(q.sum())
            if isum != INTER_REMAP_COEF_SCALE:
                diff = isum - INTER_REMAP_COEF_SCALE
                Mk = (2, 2)
                mk = (2, 2)
                for k1 in (2, 3):
                    for k2 in (2, 3):
                        if q[k1, k2] < q[mk]:
                            mk = (k1, k2)
                        elif q[k1, k2] > q[Mk]:
                            Mk = (k1, k2)
                if diff < 0:
                    q[Mk] -= diff
                else:
                    q[mk] -= diff
            out[fy * INTER_TAB_SIZE + fx] = q.reshape(16).astype(np.int16)
    return out


def cubic_table_f32() -> np.ndarray:
    """float32 [1024,16] outer-product weights used for non-u8 images."""
    t1 = cubic_coeffs_1d()
    out = np.zeros((INTER_TAB_SIZE * INTER_TAB_SIZE, 16), np.float32)
    for fy in range(INTER_TAB_SIZE):
        for fx in range(INTER_TAB_SIZE):
            out[fy * INTER_TAB_SIZE + fx] = (t1[fy][:, None] * t1[fx][None, :]).astype(np.float32).reshape(16)
    return out


_TAB_I16 = None
_TAB_F32 = None


def _tables():
    global _TAB_I16, _TAB_F32
    if _TAB_I16 is None:
        _TAB_I16 = cubic_table_i16()
        _TAB_F32 = cubic_table_f32()
    return _TAB_I16, _TAB_F32


def _fixed_point_coords(map_x: np.ndarray, map_y: np.ndarray):
    """cv2.remap's float-map -> fixed-point conversion: cvRound(v*32) (half to
    even; out-of-int-range -> INT_MIN like cvtss2si), integer part saturated
    to int16, 5 fractional bits."""
    def q(v):
        s = v.astype(np.float32) * np.float32(INTER_TAB_SIZE)
        r = np.rint(s.astype(np.float64))
        bad = ~np.isfinite(r) | (r >= 2147483648.0) | (r < -2147483648.0)
        r = np.where(bad, -2147483648.0, r).astype(np.int64)
        return r
    sx = q(map_x)
    sy = q(map_y)
    ix = np.clip(sx >> INTER_BITS, -32768, 32767)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767)
    fx = sx & (INTER_TAB_SIZE - 1)
    fy = sy & (INTER_TAB_SIZE - 1)
    return ix, iy, (fy * INTER_TAB_SIZE + fx)


def remap_cubic(src: np.ndarray, map_x: np.ndarray, map_y: np.ndarray) -> np.ndarray:
    """cv2.remap(src, map_x, map_y, INTER_CUBIC, BORDER_CONSTANT, 0) restated.

    uint8: integer arithmetic, bit-exact.  float32: same 1/32-pixel coordinate
    quantisation with fp32 weights (matches cv2 to ~1e-7 relative)."""
    tab_i, tab_f = _tables()
    squeeze = src.ndim == 2
    img = src[:, :, None] if squeeze else src
    H, W, C = img.shape
    h, w = map_x.shape
    ix, iy, fidx = _fixed_point_coords(map_x, map_y)
    sx0 = ix - 1
    sy0 = iy - 1
    if img.dtype == np.uint8:
        acc = np.zeros((h, w, C), np.int64)
        wt = tab_i[fidx].astype(np.int64)          # [h,w,16]
        imgw = img.astype(np.int64)
    else:
        acc = np.zeros((h, w, C), np.float32)
        wt = tab_f[fidx]
        imgw = img.astype(np.float32)
    for ky in range(4):
        yy = sy0 + ky
        vy = (yy >= 0) & (yy < H)
        yc = np.clip(yy, 0, H - 1)
        for kx in range(4):
            xx = sx0 + kx
            valid = vy & (xx >= 0) & (xx < W)
            xc = np.clip(xx, 0, W - 1)
            px = imgw[yc, xc]                       # [h,w,C]
            wk = wt[:, :, ky * 4 + kx]
            if img.dtype == np.uint8:
                acc += np.where(valid, wk, 0)[:, :, None] * px
            else:
                acc = acc + (np.where(valid, wk, np.float32(0))[:, :, None] * px).astype(np.float32)
    if img.dtype == np.uint8:
        out = np.clip((acc + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS, 0, 255).astype(np.uint8)
    else:
        out = acc.astype(src.dtype)
    return out[:, :, 0] if squeeze else out


def maps_pdcnet(flow: np.ndarray):
    """map = float32(float64 meshgrid + flow)   (pdcnet_of.py:35-40)."""
    h, w = flow.shape[:2]
    X, Y = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
    return (X + flow[:, :, 0]).astype(np.float32), (Y + flow[:, :, 1]).astype(np.float32)


def maps_raft(flow: np.ndarray):
    """map = -flow + arange in float32 arithmetic   (ofgen.py:38-41)."""
    h, w = flow.shape[:2]
    m = -flow.astype(np.float32)
    m = m.copy()
    m[:, :, 0] += np.arange(w)
    m[:, :, 1] += np.arange(h)[:, None]
    return np.ascontiguousarray(m[:, :, 0]), np.ascontiguousarray(m[:, :, 1])


def warp_frame_pdcnet(frame: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """pdcnet_of.warp_frame (pdcnet_of.py:34-42): sample at x + flow."""
    mx, my = maps_pdcnet(flow)
    return remap_cubic(frame, mx, my)


def warp_frame_raft(frame: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """ofgen.warp_frame (ofgen.py:37-43): sample at x - flow."""
    mx, my = maps_raft(flow)
    return remap_cubic(frame, mx, my)


def warp_bilinear(img: np.ndarray, flow: np.ndarray, sign: float = 1.0) -> np.ndarray:
    """Backward bilinear warp with zero padding at exact pixel coordinates
    x + sign*flow: the semantics of RAFT's bilinear_sampler
    (RAFT/core/utils/utils.py:57-71 = grid_sample(align_corners=True,
    padding_mode='zeros')) without the normalise/unnormalise round trip.
    img: [H,W] or [H,W,C] float32 or uint8 (uint8 output = rint, clipped)."""
    squeeze = img.ndim == 2
    im = img[:, :, None] if squeeze else img
    H, W, C = im.shape
    f = im.astype(np.float32)
    h, w = flow.shape[:2]
    xs = (np.arange(w, dtype=np.float32)[None, :] + np.float32(sign) * flow[:, :, 0].astype(np.float32)).astype(np.float32)
    ys = (np.arange(h, dtype=np.float32)[:, None] + np.float32(sign) * flow[:, :, 1].astype(np.float32)).astype(np.float32)
    x0 = np.floor(xs)
    y0 = np.floor(ys)
    ax = (xs - x0).astype(np.float32)
    ay = (ys - y0).astype(np.float32)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)

    def tap(yy, xx):
        valid = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = f[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(valid[:, :, None], v, np.float32(0))

    one = np.float32(1)
    w00 = ((one - ax) * (one - ay))[:, :, None]
    w01 = (ax * (one - ay))[:, :, None]
    w10 = ((one - ax) * ay)[:, :, None]
    w11 = (ax * ay)[:, :, None]
    out = (tap(y0, x0) * w00 + tap(y0, x0 + 1) * w01 + tap(y0 + 1, x0) * w10 + tap(y0 + 1, x0 + 1) * w11).astype(np.float32)
    if img.dtype == np.uint8:
        out = np.clip(np.rint(out), 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def warp_frame_latent(latent_chw: np.ndarray, flow: np.ndarray, sign: float = 1.0) -> np.ndarray:
    """warp_frame_latent (pdcnet_of.py:19-32): cubic-resize the [C,h,w] latent
    to the flow's size, cubic remap, cubic-resize back.  cv2.resize is used
    directly (it is what the reference calls); only the remap is restated."""
    import cv2
    lat = np.ascontiguousarray(np.transpose(latent_chw, (1, 2, 0)))
    lh, lw = lat.shape[:2]
    h, w = flow.shape[:2]
    big = cv2.resize(lat, (w, h), interpolation=cv2.INTER_CUBIC)
    if sign > 0:
        mx, my = maps_pdcnet(flow)
    else:
        mx, my = maps_raft(flow)
    rem = remap_cubic(big, mx, my)
    small = cv2.resize(rem, (lw, lh), interpolation=cv2.INTER_CUBIC)
    return np.ascontiguousarray(np.transpose(small, (2, 0, 1)))


# ----------------------------------------------------------------------------- cv2.resize(INTER_CUBIC), float32
def _cubic_coeffs_f32(x: np.ndarray) -> np.ndarray:
    """OpenCV's interpolateCubic (imgproc/resize.cpp), A = -0.75, evaluated in float32; returns [..., 4]."""
    A = np.float32(-0.75)
    x = x.astype(np.float32)
    one = np.float32(1.0)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def _resize_axis_table(n_src: int, n_dst: int):
    """Per destination index: the four source indices (replicated at the borders) and the float32 weights of
    cv::resize's INTER_CUBIC: f = float((d + 0.5) * scale - 0.5) with scale = 1 / (n_dst / n_src) in double, s = floor(f)."""
    scale = 1.0 / (float(n_dst) / float(n_src))
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    frac = (f - s.astype(np.float32)).astype(np.float32)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, n_src - 1)
    return idx, _cubic_coeffs_f32(frac)


def resize_cubic_f32(img: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.resize(img, (w, h), interpolation=cv2.INTER_CUBIC) for float32 images [H,W] / [H,W,C], restated: horizontal pass
    into float32 rows (taps summed left to right), then the vertical pass; no antialiasing (INTER_CUBIC never has any)."""
    wd, hd = dsize_wh
    src = img.astype(np.float32)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    hs, ws = src.shape[:2]
    xi, xw = _resize_axis_table(ws, wd)
    yi, yw = _resize_axis_table(hs, hd)
    rows = np.zeros((hs, wd, src.shape[2]), np.float32)
    for k in range(4):
        rows = (rows + src[:, xi[:, k], :] * xw[None, :, k, None]).astype(np.float32) if k else (src[:, xi[:, k], :] * xw[None, :, k, None]).astype(np.float32)
    out = None
    for k in range(4):
        term = (rows[yi[:, k]] * yw[:, k, None, None]).astype(np.float32)
        out = term if out is None else (out + term).astype(np.float32)
    return out[:, :, 0] if squeeze else out
