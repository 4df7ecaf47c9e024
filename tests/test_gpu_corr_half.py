"""Round-2 correlation path: fp16-STORED pyramid, split / shared (key-frame) operands, auto-ranged 16-bit operands, and
the lookup that reads the fp16 pyramid -- against the NumPy oracle (oracle/corr_oracle.py, pinned to the reference's
CorrBlock by tests/test_oracle_corr.py) and, at config sizes, against a torch fp64 matmul of sampled rows (NOT the oracle;
a size-independent spot check).

Tolerances, for unit-variance features and C channels (|corr| <= S):
  fp16 operands (11-bit significand, like TF32): 4e-3 abs (see tests/test_gpu_corr.py)
  + fp16 storage of the result: S * 2^-11 (round to nearest; the TF32 convolution that consumes the lookup keeps the same
    11 bits of it)
  bf16 operands: 3.2e-2 abs.
"""
import numpy as np
import pytest
import torch

from oracle import corr_oracle as co

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _nhwc(f, dev):
    return _t(f, dev).permute(0, 2, 3, 1).contiguous()


def _tol(precision, storage, scale):
    t = {'fp16': 4e-3, 'bf16': 3.2e-2}[precision]
    return t + (scale * 2.0 ** -11 if storage == 'fp16' else 0.0)


@pytest.mark.parametrize('storage', ['fp16', 'fp32'])
@pytest.mark.parametrize('precision', ['fp16', 'bf16'])
@pytest.mark.parametrize('shape', [(1, 256, 24, 40), (2, 128, 17, 23), (1, 40, 9, 50), (1, 256, 16, 16), (2, 32, 18, 22), (1, 64, 11, 20)])
def test_split_operand_pyramid_vs_oracle(cuda, shape, precision, storage):
    """Every element of every level: partial source blocks, partial / per-level patch shapes (32x4, 16x8, 8x16), odd pooled
    widths (the half2 pair store's padding column), B > 1, C not a multiple of the K-slab."""
    from sd_animation_optical_flow_b200 import ops
    B, C, h, w = shape
    rs = np.random.RandomState(C + h)
    f1 = rs.standard_normal(shape).astype(np.float32)
    f2 = rs.standard_normal(shape).astype(np.float32)
    ref = co.corr_pyramid(f1, f2, 4)
    pyr = ops.corr_volume_pyramid(_nhwc(f1, cuda), _nhwc(f2, cuda), 4, precision, storage)
    assert pyr.buf.dtype == (torch.float16 if storage == 'fp16' else torch.float32)
    scale = float(np.abs(ref[0]).max())
    for l in range(4):
        got = pyr.level_values(l)[:, 0].cpu().numpy()
        assert got.shape == ref[l].shape
        if got.size:
            err = np.abs(got - ref[l]).max()
            assert err <= _tol(precision, storage, scale), f'{precision}/{storage} level {l}: max abs err {err} (max|corr| {scale})'
    # the lookup that reads this pyramid, vs the oracle's lookup in the exact pyramid
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing='ij')
    coords = (np.stack([xs, ys], 0)[None].repeat(B, 0) + 3 * rs.standard_normal((B, 2, h, w))).astype(np.float32)
    coords[:, :, 0, 0] = [0.0, 0.0]
    coords[:, 0, 1, 1] = -6.5
    coords[:, 1, 2, 2] = h + 7.25
    look_ref = co.corr_lookup(ref, coords, 4)
    look = ops.corr_lookup(pyr, _t(coords, cuda), 4).cpu().numpy()
    assert np.abs(look - look_ref).max() <= _tol(precision, storage, scale) + 3e-5
    nh = torch.empty((B, h, w, 324), device=cuda)
    ops.corr_lookup_nhwc(pyr, _t(coords, cuda).permute(0, 2, 3, 1).contiguous(), 4, nh)
    assert torch.equal(nh.permute(0, 3, 1, 2), _t(look, cuda))


def test_shared_key_target_equals_per_pair_targets(cuda):
    """One prepared target (batch 1) serving B pairs == B separate pairs against the same fmap2, bit for bit."""
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(3)
    B, h, w, C = 3, 24, 40, 256
    f1 = torch.randn((B, h, w, C), generator=g, device=cuda)
    key = torch.randn((1, h, w, C), generator=g, device=cuda)
    tgt = ops.CorrTarget(key, 4, 'fp16')
    shared = ops.corr_volume_pyramid(f1, None, 4, 'fp16', 'fp16', target=tgt)
    again = ops.corr_volume_pyramid(f1, key, 4, 'fp16', 'fp16')                       # fmap2 of batch 1 broadcasts too
    for b in range(B):
        single = ops.corr_volume_pyramid(f1[b:b + 1].contiguous(), key, 4, 'fp16', 'fp16')
        for l in range(4):
            rows = slice(b * h * w, (b + 1) * h * w)
            assert torch.equal(shared.level(l)[rows], single.level(l)) and torch.equal(shared.factor, single.factor)
            assert torch.equal(again.level(l)[rows], single.level(l))
    with pytest.raises(RuntimeError):
        ops.CorrSource(f1, 'fp16').pyramid(ops.CorrTarget(torch.randn((2, h, w, C), device=cuda), 4, 'fp16'))


@pytest.mark.parametrize('s1,s2', [(1e3, 1e3), (1e-4, 1e-4), (3e4, 1e-5), (1.0, 1.0)])
def test_fp16_operands_are_auto_ranged(cuda, s1, s2):
    """Feature magnitudes far from 1: with round 1's fixed 1/16 pre-scale, x1e-4 features fell into fp16
    subnormals (1 % relative error per element) and large ones saturated at 65504.  With the per-tensor power-of-two scales
    the RELATIVE error is the unit-scale one for any magnitude -- for the fp16-STORED pyramid too, whose stored values
    are the raw accumulators (|acc| < 2^15 by construction) with the factor back to correlation units in the header."""
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(5)
    h, w, C = 24, 32, 256
    f1 = torch.randn((1, h, w, C), generator=g, device=cuda) * s1
    f2 = torch.randn((1, h, w, C), generator=g, device=cuda) * s2
    exact = (f1.double().reshape(-1, C) @ f2.double().reshape(-1, C).t() / 16.0)
    big = float(exact.abs().max())
    for storage in ('fp32', 'fp16'):
        got = ops.corr_volume_pyramid(f1, f2, 1, 'fp16', storage).level_values(0).double().reshape(h * w, h * w)
        rel = float((got - exact).abs().max()) / big
        print(f'scales {s1:g} x {s2:g}, storage {storage}: max|corr| {big:.3g}, max rel err {rel:.2e}')
        assert rel <= (8e-4 if storage == 'fp32' else 8e-4 + 2.0 ** -11)
    # all-zero and non-finite inputs must not poison the scale
    z = ops.corr_volume_pyramid(torch.zeros_like(f1), f2, 1, 'fp16', 'fp32').level(0)
    assert float(z.abs().max()) == 0.0
    f1n = f1.clone()
    f1n[0, 0, 0, 0] = float('inf')
    bad = ops.corr_volume_pyramid(f1n, f2, 1, 'fp16', 'fp32').level(0).reshape(h * w, h * w)
    assert torch.isfinite(bad[1:]).all()      # only the row of the poisoned source pixel is affected


@pytest.mark.parametrize('hw', [(96, 64), (90, 160)])
def test_config_size_fp16_pyramid_spot_check(cuda, hw):
    """Config 2 / config 5 operator sizes with the product defaults (fp16 operands, fp16 storage): 64 sampled source rows of
    level 0 against a torch fp64 matmul (a size-independent check, not the oracle), pooled levels against avg_pool2d of the
    fp64 rows, and the lookup at integer coordinates reading the volume back."""
    import torch.nn.functional as F
    from sd_animation_optical_flow_b200 import ops
    from sd_animation_optical_flow_b200.raft import coords_grid
    h, w = hw
    g = torch.Generator(device=cuda).manual_seed(h)
    f1 = torch.randn((1, h, w, 256), generator=g, device=cuda)
    f2 = torch.randn((1, h, w, 256), generator=g, device=cuda)
    pyr = ops.corr_volume_pyramid(f1, f2, 4, 'fp16', 'fp16')
    rows = torch.linspace(0, h * w - 1, 64, device=cuda).long()
    exact = (f1.double().reshape(-1, 256)[rows] @ f2.double().reshape(-1, 256).t() / 16.0).reshape(64, 1, h, w)
    scale = float(exact.abs().max())
    cur = exact
    for l in range(4):
        if l:
            cur = F.avg_pool2d(cur, 2, stride=2)
        got = pyr.level_values(l)[rows].double()
        assert got.shape == cur.shape
        assert float((got - cur).abs().max()) <= _tol('fp16', 'fp16', scale)
    look = ops.corr_lookup(pyr, coords_grid(1, h, w, cuda), 4)
    # centre tap of level 0 (channel 9*4+4 = 40) at integer coords = the volume's diagonal
    diag = (pyr.level(0)[:, 0].reshape(h * w, h * w).diagonal().float() * pyr.factor).reshape(h, w)
    assert torch.equal(look[0, 40], diag)
