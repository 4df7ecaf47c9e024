"""Oracle for the all-pairs correlation volume, its pyramid and the windowed
lookup.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  NumPy fp32 throughout.
Pinned against the reference's own `CorrBlock` (tests/golden/corr_*.npz made by
oracle/make_golden.py from /root/reference/RAFT/core/corr.py).
"""
from __future__ import annotations

import numpy as np


def corr_volume(fmap1: np.ndarray, fmap2: np.ndarray) -> np.ndarray:
    """CorrBlock.corr (RAFT/core/corr.py:52-60).
    fmap1 [B,C,h1,w1], fmap2 [B,C,h2,w2] -> [B, h1*w1, h2, w2] fp32,
    corr[b,i,j] = <fmap1[b,:,i], fmap2[b,:,j]> / sqrt(C)."""
    B, C, h1, w1 = fmap1.shape
    _, _, h2, w2 = fmap2.shape
    a = fmap1.reshape(B, C, h1 * w1).astype(np.float32)
    b = fmap2.reshape(B, C, h2 * w2).astype(np.float32)
    out = np.matmul(np.transpose(a, (0, 2, 1)), b).astype(np.float32)
    out = out / np.sqrt(np.float32(C))
    return out.reshape(B, h1 * w1, h2, w2).astype(np.float32)


def avg_pool2(x: np.ndarray) -> np.ndarray:
    """F.avg_pool2d(x, 2, stride=2) on the last two dims, floor mode
    (RAFT/core/corr.py:25-27).  Sum order follows ATen: ((a00+a01)+a10)+a11."""
    h, w = x.shape[-2:]
    ho, wo = h // 2, w // 2
    x = x[..., : 2 * ho, : 2 * wo].astype(np.float32)
    s = (x[..., 0::2, 0::2] + x[..., 0::2, 1::2]).astype(np.float32)
    s = (s + x[..., 1::2, 0::2]).astype(np.float32)
    s = (s + x[..., 1::2, 1::2]).astype(np.float32)
    return (s * np.float32(0.25)).astype(np.float32)


def corr_pyramid(fmap1: np.ndarray, fmap2: np.ndarray, num_levels: int = 4):
    """CorrBlock.__init__ (RAFT/core/corr.py:12-27): list of [B*N1, h_l, w_l]."""
    vol = corr_volume(fmap1, fmap2)
    B, N1, h2, w2 = vol.shape
    lv = vol.reshape(B * N1, h2, w2)
    pyr = [lv]
    for _ in range(num_levels - 1):
        lv = avg_pool2(lv)
        pyr.append(lv)
    return pyr


def _bilinear_zero(maps: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """maps [P,H,W]; xs, ys [P,K] pixel coords -> [P,K]; zeros outside
    (bilinear_sampler, RAFT/core/utils/utils.py:57-71, align_corners=True)."""
    P, H, W = maps.shape
    x0 = np.floor(xs)
    y0 = np.floor(ys)
    ax = (xs - x0).astype(np.float32)
    ay = (ys - y0).astype(np.float32)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    pi = np.arange(P)[:, None]

    def tap(yy, xx):
        valid = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = maps[pi, np.clip(yy, 0, max(H - 1, 0)), np.clip(xx, 0, max(W - 1, 0))]
        return np.where(valid, v, np.float32(0)).astype(np.float32)

    one = np.float32(1)
    return (tap(y0, x0) * ((one - ax) * (one - ay)) + tap(y0, x0 + 1) * (ax * (one - ay))
            + tap(y0 + 1, x0) * ((one - ax) * ay) + tap(y0 + 1, x0 + 1) * (ax * ay)).astype(np.float32)


def corr_lookup(pyramid, coords: np.ndarray, radius: int = 4) -> np.ndarray:
    """CorrBlock.__call__ (RAFT/core/corr.py:29-50).
    pyramid: list of [B*N1,h_l,w_l]; coords [B,2,h1,w1] (x,y) -> [B, L*(2r+1)^2, h1, w1].
    Channel order is x-major: k = lvl*(2r+1)^2 + (2r+1)*ix + iy, because the
    reference adds meshgrid(dy,dx) to (x,y) coordinates (corr.py:37-43)."""
    B, _, h1, w1 = coords.shape
    r = radius
    d = 2 * r + 1
    P = B * h1 * w1
    cx = np.transpose(coords, (0, 2, 3, 1)).reshape(P, 2).astype(np.float32)
    offs = np.arange(-r, r + 1, dtype=np.float32)
    outs = []
    for lvl, maps in enumerate(pyramid):
        cen = (cx / np.float32(2 ** lvl)).astype(np.float32)
        # sample [ix, iy] at (x + offs[ix], y + offs[iy])
        xs = (cen[:, 0][:, None, None] + offs[None, :, None]).astype(np.float32)
        ys = (cen[:, 1][:, None, None] + offs[None, None, :]).astype(np.float32)
        xs = np.broadcast_to(xs, (P, d, d)).reshape(P, d * d)
        ys = np.broadcast_to(ys, (P, d, d)).reshape(P, d * d)
        outs.append(_bilinear_zero(maps, xs, ys))
    out = np.concatenate(outs, axis=1).reshape(B, h1, w1, -1)
    return np.ascontiguousarray(np.transpose(out, (0, 3, 1, 2))).astype(np.float32)


def alt_corr_forward(fmap1: np.ndarray, fmap2: np.ndarray, coords: np.ndarray, radius: int) -> np.ndarray:
    """corr_forward_kernel (RAFT/alt_cuda_corr/correlation_kernel.cu:18-119) as
    called from AlternateCorrBlock (RAFT/core/corr.py:74-91).
    fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C] channels-last, coords [B,N,H1,W1,2]
    (x,y in fmap2 pixels) -> corr [B,N,(2r+1)^2,H1,W1], UNNORMALISED.
    For each of the (2r+2)^2 integer taps around floor(coords)-r the kernel
    takes the C-long dot product and scatters it to <=4 outputs with weights
    nw=dy*dx, ne=dy*(1-dx), sw=(1-dy)*dx, se=(1-dy)*(1-dx) (:92-114); output
    index = iy + (2r+1)*ix.  Equivalent gather form used here:
      out[ix,iy] = sum_{a,b in {0,1}} wy_a*wx_b * dot(tap[iy+a, ix+b])."""
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    N = coords.shape[1]
    r = radius
    d = 2 * r + 1
    out = np.zeros((B, N, d * d, H1, W1), np.float32)
    f1 = fmap1.astype(np.float32)
    f2 = fmap2.astype(np.float32)
    for b in range(B):
        for n in range(N):
            x = coords[b, n, :, :, 0].astype(np.float32)
            y = coords[b, n, :, :, 1].astype(np.float32)
            fx = np.floor(x)
            fy = np.floor(y)
            dx = (x - fx).astype(np.float32)
            dy = (y - fy).astype(np.float32)
            x0 = fx.astype(np.int64) - r
            y0 = fy.astype(np.int64) - r
            dots = np.zeros((d + 1, d + 1, H1, W1), np.float32)  # [iy, ix]
            for iy in range(d + 1):
                for ix in range(d + 1):
                    yy = y0 + iy
                    xx = x0 + ix
                    valid = (yy >= 0) & (yy < H2) & (xx >= 0) & (xx < W2)
                    g = f2[b, np.clip(yy, 0, H2 - 1), np.clip(xx, 0, W2 - 1)]     # [H1,W1,C]
                    s = np.einsum('hwc,hwc->hw', f1[b], g).astype(np.float32)
                    dots[iy, ix] = np.where(valid, s, np.float32(0))
            one = np.float32(1)
            for ix in range(d):
                for iy in range(d):
                    v = (dots[iy, ix] * ((one - dy) * (one - dx)) + dots[iy, ix + 1] * ((one - dy) * dx)
                         + dots[iy + 1, ix] * (dy * (one - dx)) + dots[iy + 1, ix + 1] * (dy * dx))
                    out[b, n, iy + d * ix] = v.astype(np.float32)
    return out


def alternate_corr_block(fmap1: np.ndarray, fmap2: np.ndarray, coords: np.ndarray,
                         num_levels: int = 4, radius: int = 4) -> np.ndarray:
    """AlternateCorrBlock (RAFT/core/corr.py:63-91): pools fmap2, calls K1 per
    level with coords/2^i, stacks, divides by sqrt(C).  fmaps [B,C,h,w]."""
    B, C, h, w = fmap1.shape
    f1 = np.ascontiguousarray(np.transpose(fmap1, (0, 2, 3, 1)))
    cur2 = fmap2.astype(np.float32)
    c = np.transpose(coords, (0, 2, 3, 1)).astype(np.float32)
    outs = []
    for i in range(num_levels):
        f2 = np.ascontiguousarray(np.transpose(cur2, (0, 2, 3, 1)))
        ci = (c / np.float32(2 ** i)).reshape(B, 1, h, w, 2).astype(np.float32)
        outs.append(alt_corr_forward(f1, f2, ci, radius)[:, 0])
        cur2 = avg_pool2(cur2)
    out = np.stack(outs, axis=1).reshape(B, -1, h, w)
    return (out / np.sqrt(np.float32(C))).astype(np.float32)
