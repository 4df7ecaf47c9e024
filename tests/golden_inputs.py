"""Seeded input generators shared by oracle/make_golden.py (which stores the REFERENCE's
outputs for them under tests/golden/) and the tests.  Pure NumPy with the frozen legacy
RandomState streams, so the inputs are identical on every box."""
from __future__ import annotations

import numpy as np

# name -> (B, C, h, w) of the feature maps
CORR_CASES = {
    'even': (1, 64, 16, 24),     # every level even-sized
    'odd': (2, 32, 18, 22),      # 18x22 -> 9x11 -> 4x5 -> 2x2 : floor pooling drops rows/cols
}


def PYRAMID_ROWS(n_rows: int):
    """Source-pixel rows of each pyramid level kept in the fixture (the full volume is too big)."""
    return np.unique(np.linspace(0, n_rows - 1, 24).astype(np.int64))


def corr_inputs(name: str):
    B, C, h, w = CORR_CASES[name]
    rs = np.random.RandomState(1234 + len(name))
    f1 = rs.standard_normal((B, C, h, w)).astype(np.float32)
    f2 = rs.standard_normal((B, C, h, w)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing='ij')
    grid = np.stack([xs, ys], 0).astype(np.float32)[None].repeat(B, 0)
    coords = grid + 3.0 * rs.standard_normal((B, 2, h, w)).astype(np.float32)
    coords[:, :, 0, 0] = grid[:, :, 0, 0]              # exactly integer coordinates
    coords[:, 0, 1, 1] = -7.25                          # window partly / fully outside
    coords[:, 1, 2, 2] = h + 9.5
    return f1, f2, np.ascontiguousarray(coords.astype(np.float32))


def _blur(img: np.ndarray, sigma: float) -> np.ndarray:
    r = int(3 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    out = img.astype(np.float64)
    for ax in (0, 1):
        pad = [(0, 0)] * out.ndim
        pad[ax] = (r, r)
        p = np.pad(out, pad, mode='reflect')
        out = sum(k[i] * np.take(p, range(i, i + img.shape[ax]), axis=ax) for i in range(2 * r + 1))
    return out


def texture(h: int, w: int, seed: int, sigma: float = 2.0, channels: int = 3) -> np.ndarray:
    """Gaussian-blurred uniform noise stretched to 0..255, uint8 [h,w,channels]."""
    rs = np.random.RandomState(seed)
    t = _blur(rs.uniform(0, 1, (h, w, channels)), sigma)
    t = (t - t.min()) / (t.max() - t.min())
    return np.clip(np.rint(t * 255), 0, 255).astype(np.uint8)


def shifted_pair(h: int, w: int, seed: int, dx: int = 4, dy: int = -3):
    """Frame 2 = frame 1 shifted by (+dx, +dy) px: crops of one larger canvas."""
    m = 16
    canvas = texture(h + 2 * m, w + 2 * m, seed)
    f1 = canvas[m:m + h, m:m + w]
    f2 = canvas[m - dy:m - dy + h, m - dx:m - dx + w]
    return np.ascontiguousarray(f1), np.ascontiguousarray(f2)


RAFT_CASES = {
    'basic': dict(small=False, seed=0, iters=6, hw=(128, 160)),
    'basic_pad': dict(small=False, seed=1, iters=4, hw=(132, 150)),   # exercises InputPadder
    'small': dict(small=True, seed=2, iters=4, hw=(128, 144)),
}


def raft_inputs(name: str):
    cfg = RAFT_CASES[name]
    h, w = cfg['hw']
    return shifted_pair(h, w, 100 + cfg['seed'])   # RGB uint8 [H,W,3] x2


# the sizes bench.py times (configs[1..3]: 768x512; configs[4]: 720x1280), iters as in ofgen.py:77; 'S' is exactly the pair
# of bench.py's rank 0 (bench.synthetic_pair(0)) with the bench weights (seed 0).  Name-seeded random weights make the
# update block an amplifier: the flow runs away to ~100 px in 20 iterations (nothing like a trained model), so every case
# also exists in a `calm` variant whose flow-head output convolution is scaled by `fh_scale` (flow stays a few px, the
# regime a trained checkpoint works in and the one in which the correlation lookups stay inside the volume).
RAFT_FULL_CASES = {
    'S': dict(seed=0, iters=20, hw=(768, 512), pair_seed=1000, shift=(4, -3), fh_scale=1.0),
    'S_calm': dict(seed=0, iters=20, hw=(768, 512), pair_seed=1000, shift=(4, -3), fh_scale=0.02),
    'L': dict(seed=0, iters=20, hw=(720, 1280), pair_seed=1077, shift=(5, -2), fh_scale=1.0),
    'L_calm': dict(seed=0, iters=20, hw=(720, 1280), pair_seed=1077, shift=(5, -2), fh_scale=0.02),
}


def raft_full_inputs(name: str):
    cfg = RAFT_FULL_CASES[name]
    h, w = cfg['hw']
    return shifted_pair(h, w, cfg['pair_seed'], dx=cfg['shift'][0], dy=cfg['shift'][1])


def raft_full_weights(model, name: str):
    """Name-seeded weights of a full-size case on `model` (the reference RAFT or this package's)."""
    from sd_animation_optical_flow_b200.raft import fill_weights_by_name
    cfg = RAFT_FULL_CASES[name]
    return fill_weights_by_name(model, cfg['seed'], flow_head_scale=cfg['fh_scale'])


def full_lattice(H: int, W: int):
    """Pixels of flow_up kept in the fixture: one per 8x8 block, its phase inside the block varying from block to block
    so that all 64 positions of the convex-upsampling stencil are sampled."""
    i, j = np.arange(H // 8), np.arange(W // 8)
    return 8 * i + (3 * i) % 8, 8 * j + (5 * j) % 8


WARP_CASES = ('small_u8', 'wild_u8', 'gray_f32', 'rgb_f32')


def warp_inputs(name: str):
    rs = np.random.RandomState(77 + WARP_CASES.index(name))
    if name == 'small_u8':
        img = rs.randint(0, 256, (48, 64, 3)).astype(np.uint8)
        flow = (4.0 * rs.standard_normal((48, 64, 2))).astype(np.float32)
    elif name == 'wild_u8':
        # borders, far out-of-image, NaN / inf / huge coordinates, exact integers and exact 1/64 ties
        img = rs.randint(0, 256, (37, 53, 3)).astype(np.uint8)
        flow = (15.0 * rs.standard_normal((37, 53, 2))).astype(np.float32)
        flow[0, :8] = [1e9, -1e9]
        flow[1, :4, 0] = np.nan
        flow[2, :4, 1] = np.inf
        flow[3, :4] = -np.inf
        flow[4, :8] = [40000.0, 3.0]
        flow[5, :16, 0] = np.arange(16) / 64.0
        flow[5, :16, 1] = -np.arange(16) / 64.0
        flow[6, :8] = 0.0
        flow[7, :8] = [-1e12, 1e12]
    elif name == 'gray_f32':
        img = rs.standard_normal((40, 56)).astype(np.float32)
        flow = (5.0 * rs.standard_normal((40, 56, 2))).astype(np.float32)
    else:
        img = (255 * rs.uniform(0, 1, (32, 40, 3))).astype(np.float32)
        flow = (6.0 * rs.standard_normal((32, 40, 2))).astype(np.float32)
    return img, flow


def mask_inputs():
    """(confidence f32 [H,W], log_confidence f32 [H,W], image u8 [H,W,3], raw u8, warped u8)."""
    rs = np.random.RandomState(4242)
    H, W = 45, 61                                       # odd sizes: tile edges + unaligned rows
    conf = _blur(rs.uniform(0, 1, (H, W)), 1.5)
    conf = ((conf - conf.min()) / (conf.max() - conf.min())).astype(np.float32)
    logc = np.log(np.maximum(conf, 1e-6)).astype(np.float32)
    img = texture(H, W, 9, sigma=1.0)
    raw = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    warped = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    return conf, logc, img, raw, warped


def greedy_inputs():
    """(flow_mat f32 [n,1,H,W,3], frames u8 [n,H,W,3], thres)."""
    rs = np.random.RandomState(99)
    n, H, W = 4, 40, 52
    fm = np.zeros((n, 1, H, W, 3), np.float32)
    fm[..., 0:2] = (3.0 * rs.standard_normal((n, 1, H, W, 2))).astype(np.float32)
    for s in range(n):
        c = _blur(rs.uniform(0, 1, (H, W)), 3.0)
        fm[s, 0, :, :, 2] = ((c - c.min()) / (c.max() - c.min())).astype(np.float32)
    frames = rs.randint(0, 256, (n, H, W, 3)).astype(np.uint8)
    return fm, frames, 0.55
