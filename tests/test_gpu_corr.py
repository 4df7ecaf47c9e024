"""GPU parity of the correlation kernels through the C ABI.

Tolerances (stated per SURVEY §8d):
  fp32 (CUDA-core) volume vs oracle/golden : 2e-5 abs  (|corr| <= ~6 here; summation-order only)
  3xtf32 volume                            : 1e-5 relative to max|corr|
  tf32 volume (features rounded to tf32)   : 4e-3 abs for unit-variance features: each operand carries a
                                             2^-12 relative rounding error, so a C-term dot product / sqrt(C)
                                             has sigma ~= 4e-4 and the max over ~1e7 entries stays < 4e-3
                                             (SURVEY's bound for real features: 1e-2 abs on |corr| <= 65)
  fp16 volume (the default fast path)      : same as tf32 (fp16 has the same 11-bit significand)
  bf16 volume                              : 3.2e-2 abs (8x the tf32 operand error)
  fp16/bf16 pooled levels                  : built from avg-pooled fmap2 (linear identity, like the reference's
                                             AlternateCorrBlock), so they match avg_pool2d of level 0 to the
                                             operand rounding error, not to 1e-6
  pooled levels vs avg_pool2d of level 0   : 1e-6 abs (same summation order as ATen)
  lookup vs reference CorrBlock            : 3e-5 abs on top of the volume error
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import corr_oracle as co
from tests import golden_inputs as gi

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _nhwc(f, dev):
    return _t(f, dev).permute(0, 2, 3, 1).contiguous()


def _vol_tol(precision, scale):
    return {'fp32': 2e-5, '3xtf32': 1e-5 * scale + 2e-5, 'tf32': 4e-3, 'fp16': 4e-3, 'bf16': 3.2e-2}[precision]


@pytest.mark.parametrize('precision', ['fp32', '3xtf32', 'tf32', 'fp16', 'bf16'])
@pytest.mark.parametrize('name', list(gi.CORR_CASES))
def test_volume_pyramid_vs_reference_golden(cuda, golden, name, precision):
    from sd_animation_optical_flow_b200 import ops
    f1, f2, _ = gi.corr_inputs(name)
    pyr = ops.corr_volume_pyramid(_nhwc(f1, cuda), _nhwc(f2, cuda), 4, precision)
    scale = float(np.abs(golden['corr'][f'{name}_pyr0']).max())
    for l in range(4):
        lv = pyr.level(l)[:, 0].cpu().numpy()
        rows = gi.PYRAMID_ROWS(lv.shape[0])
        ref = golden['corr'][f'{name}_pyr{l}']
        assert lv[rows].shape == ref.shape
        err = np.abs(lv[rows] - ref).max()
        assert err <= _vol_tol(precision, scale), f'level {l}: max abs err {err} (max|corr| {scale})'


@pytest.mark.parametrize('precision', ['fp32', '3xtf32', 'tf32', 'fp16', 'bf16'])
@pytest.mark.parametrize('shape', [(1, 256, 24, 40), (2, 128, 17, 23), (1, 36, 9, 50), (1, 256, 16, 16)])
def test_volume_full_check_vs_oracle(cuda, shape, precision):
    """Every element of every level, incl. partial TMA tiles (h, w not multiples of 8/32), B > 1,
    C that is not a multiple of the K-slab, and h1*w1 that is not a multiple of 128."""
    from sd_animation_optical_flow_b200 import ops
    B, C, h, w = shape
    rs = np.random.RandomState(C + h)
    f1 = rs.standard_normal(shape).astype(np.float32)
    f2 = rs.standard_normal(shape).astype(np.float32)
    ref = co.corr_pyramid(f1, f2, 4)
    pyr = ops.corr_volume_pyramid(_nhwc(f1, cuda), _nhwc(f2, cuda), 4, precision)
    scale = float(np.abs(ref[0]).max())
    for l in range(4):
        got = pyr.level(l)[:, 0].cpu().numpy()
        assert got.shape == ref[l].shape
        if got.size:
            err = np.abs(got - ref[l]).max()
            assert err <= _vol_tol(precision, scale), f'{precision} level {l}: max abs err {err} (max|corr| {scale})'


@pytest.mark.parametrize('precision', ['fp16', 'tf32', 'bf16'])
def test_volume_config2_size_properties(cuda, precision):
    """SURVEY config 2 operator size (N=6144, C=256): properties that need no CPU oracle.
    (1) pooled levels == avg_pool2d chain of level 0, (2) tensor-core result within tolerance of the
    CUDA-core fp32 result, (3) corr(f1,f2)[i,j] == corr(f2,f1)[j,i]."""
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(0)
    f1 = torch.randn((1, 96, 64, 256), generator=g, device=cuda)
    f2 = torch.randn((1, 96, 64, 256), generator=g, device=cuda)
    pyr = ops.corr_volume_pyramid(f1, f2, 4, precision)
    l0 = pyr.level(0)
    cur = l0
    pool_tol = 1e-6 if precision == 'tf32' else 2 * _vol_tol(precision, 1.0)
    for l in range(1, 4):
        cur = F.avg_pool2d(cur, 2, stride=2)
        assert torch.allclose(pyr.level(l), cur, rtol=0, atol=pool_tol), f'level {l}'
    exact = ops.corr_volume_pyramid(f1, f2, 1, 'fp32').level(0)
    scale = float(exact.abs().max())
    err = float((l0 - exact).abs().max())
    assert err <= _vol_tol(precision, scale), f'max abs err {err}, max|corr| {scale}'
    swapped = ops.corr_volume_pyramid(f2, f1, 1, precision).level(0)
    a = l0.reshape(6144, 6144)
    b = swapped.reshape(6144, 6144).t()
    assert torch.equal(a, b) or float((a - b).abs().max()) <= 1e-5 * scale


def test_volume_config5_size_runs_and_pools(cuda):
    """720x1280 (90x160 features): floor pooling 90 -> 45 -> 22 -> 11; 1.10 GB pyramid."""
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(1)
    f1 = torch.randn((1, 90, 160, 256), generator=g, device=cuda)
    f2 = torch.randn((1, 90, 160, 256), generator=g, device=cuda)
    pyr = ops.corr_volume_pyramid(f1, f2, 4, 'fp16')
    assert [tuple(pyr.level(l).shape[-2:]) for l in range(4)] == [(90, 160), (45, 80), (22, 40), (11, 20)]
    cur = pyr.level(0)
    for l in range(1, 4):
        cur = F.avg_pool2d(cur, 2, stride=2)
        assert torch.allclose(pyr.level(l), cur, rtol=0, atol=2 * _vol_tol('fp16', 1.0))
    exact_pyr = ops.corr_volume_pyramid(f1, f2, 4, 'tf32')     # streaming kernel: pooled levels == avg_pool2d chain
    cur = exact_pyr.level(0)
    for l in range(1, 4):
        cur = F.avg_pool2d(cur, 2, stride=2)
        assert torch.allclose(exact_pyr.level(l), cur, rtol=0, atol=1e-6)
    # spot-check 64 rows against a direct fp32 matmul
    rows = torch.randint(0, 14400, (64,), device=cuda)
    a = f1.reshape(14400, 256)[rows]
    ref = (a.double() @ f2.reshape(14400, 256).double().t() / 16.0).float()
    got = pyr.level(0).reshape(14400, 14400)[rows]
    assert float((got - ref).abs().max()) <= _vol_tol('tf32', float(ref.abs().max()))


@pytest.mark.parametrize('name', list(gi.CORR_CASES))
def test_corrblock_protocol_vs_reference_golden(cuda, golden, name):
    """corr_fn = CorrBlock(fmap1, fmap2, radius=4); corr_fn(coords) as RAFT.forward uses it (raft.py:104-124)."""
    from sd_animation_optical_flow_b200.corr import AlternateCorrBlock, CorrBlock
    f1, f2, coords = gi.corr_inputs(name)
    ref = golden['corr'][f'{name}_lookup']
    fn = CorrBlock(_t(f1, cuda), _t(f2, cuda), num_levels=4, radius=4, precision='fp32')
    out = fn(_t(coords, cuda))
    assert out.shape == ref.shape and out.is_contiguous() and out.dtype == torch.float32
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=3e-5)
    for prec, tol in (('3xtf32', 5e-5), ('tf32', 4e-3), ('fp16', 4e-3), ('bf16', 6e-2)):
        o = CorrBlock(_t(f1, cuda), _t(f2, cuda), radius=4, precision=prec)(_t(coords, cuda))
        assert float(np.abs(o.cpu().numpy() - ref).max()) <= tol, prec
    alt = AlternateCorrBlock(_t(f1, cuda), _t(f2, cuda), num_levels=4, radius=4)(_t(coords, cuda))
    np.testing.assert_allclose(alt.cpu().numpy(), ref, rtol=0, atol=5e-5)
    # reference attribute: list of [B*h*w, 1, h_l, w_l]
    B, C, h, w = f1.shape
    assert [tuple(t.shape) for t in fn.corr_pyramid] == [(B * h * w, 1, h >> l, w >> l) for l in range(4)]


@pytest.mark.parametrize('radius,levels', [(3, 4), (4, 2), (2, 3), (0, 1)])
def test_lookup_generic_radius_levels(cuda, radius, levels):
    from sd_animation_optical_flow_b200 import ops
    f1, f2, coords = gi.corr_inputs('odd')
    pyr = ops.corr_volume_pyramid(_nhwc(f1, cuda), _nhwc(f2, cuda), levels, 'fp32')
    out = ops.corr_lookup(pyr, _t(coords, cuda), radius).cpu().numpy()
    ref = co.corr_lookup(co.corr_pyramid(f1, f2, levels), coords, radius)
    np.testing.assert_allclose(out, ref, rtol=0, atol=3e-5)


def test_lookup_at_integer_coords_reads_the_volume(cuda):
    """Size-independent property at config-2 size: at integer coords channel 9*ix+iy of level 0 is
    volume[p, y+iy-4, x+ix-4] exactly (zero outside)."""
    from sd_animation_optical_flow_b200 import ops
    from sd_animation_optical_flow_b200.raft import coords_grid
    g = torch.Generator(device=cuda).manual_seed(2)
    f1 = torch.randn((1, 96, 64, 256), generator=g, device=cuda)
    f2 = torch.randn((1, 96, 64, 256), generator=g, device=cuda)
    pyr = ops.corr_volume_pyramid(f1, f2, 4, 'tf32')
    look = ops.corr_lookup(pyr, coords_grid(1, 96, 64, cuda), 4)
    vol = pyr.level(0).reshape(96, 64, 96, 64)
    padded = F.pad(vol, (4, 4, 4, 4))
    for ix, iy in ((0, 0), (8, 8), (4, 4), (1, 7)):
        ys = torch.arange(96, device=cuda)[:, None].expand(96, 64)
        xs = torch.arange(64, device=cuda)[None, :].expand(96, 64)
        want = padded[ys, xs, ys + iy, xs + ix]
        assert torch.equal(look[0, 9 * ix + iy], want)


def test_alt_cuda_corr_forward_dropin(cuda):
    """Same contract as the reference's pybind op (correlation.cpp:23-33): list with one
    [B,N,81,H1,W1] tensor, unnormalised; RuntimeError on non-contiguous / CPU input."""
    from sd_animation_optical_flow_b200 import alt_cuda_corr
    rs = np.random.RandomState(21)
    B, H1, W1, H2, W2, C, N, r = 2, 11, 13, 9, 15, 64, 2, 4
    f1 = rs.standard_normal((B, H1, W1, C)).astype(np.float32)
    f2 = rs.standard_normal((B, H2, W2, C)).astype(np.float32)
    coords = (rs.uniform(-3, 17, (B, N, H1, W1, 2))).astype(np.float32)
    out = alt_cuda_corr.forward(_t(f1, cuda), _t(f2, cuda), _t(coords, cuda), r)
    assert isinstance(out, list) and len(out) == 1 and tuple(out[0].shape) == (B, N, 81, H1, W1)
    ref = co.alt_corr_forward(f1, f2, coords, r)
    np.testing.assert_allclose(out[0].cpu().numpy(), ref, rtol=0, atol=2e-4)   # |corr| ~ sqrt(64)*3
    with pytest.raises(RuntimeError, match='contiguous'):
        alt_cuda_corr.forward(_t(f1, cuda).permute(0, 2, 1, 3), _t(f2, cuda), _t(coords, cuda), r)
    with pytest.raises(RuntimeError, match='CUDA'):
        alt_cuda_corr.forward(torch.from_numpy(f1), _t(f2, cuda), _t(coords, cuda), r)
    with pytest.raises(NotImplementedError):
        alt_cuda_corr.backward(None, None, None, None, r)
    # radius 3 (small model) and C = 256
    f1b = rs.standard_normal((1, 8, 8, 256)).astype(np.float32)
    f2b = rs.standard_normal((1, 8, 8, 256)).astype(np.float32)
    cb = rs.uniform(0, 8, (1, 1, 8, 8, 2)).astype(np.float32)
    o3, = alt_cuda_corr.forward(_t(f1b, cuda), _t(f2b, cuda), _t(cb, cuda), 3)
    np.testing.assert_allclose(o3.cpu().numpy(), co.alt_corr_forward(f1b, f2b, cb, 3), rtol=0, atol=5e-4)


def test_prepared_operands_reuse_key_frame(cuda):
    """Key-frame scheme: fmap2 (key frame) operands are prepared once, fmap1 changes per pair; the result must be
    bit-identical to the one-shot op."""
    from sd_animation_optical_flow_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(5)
    key = torch.randn((1, 24, 40, 256), generator=g, device=cuda)
    ops_h = ops.CorrOperands(1, 24, 40, 24, 40, 256, 4, 'fp16', cuda).prepare(fmap2_nhwc=key)
    for _ in range(2):
        cur = torch.randn((1, 24, 40, 256), generator=g, device=cuda)
        pyr = ops_h.prepare(fmap1_nhwc=cur).pyramid()
        ref = ops.corr_volume_pyramid(cur, key, 4, 'fp16')
        for l in range(4):
            assert torch.equal(pyr.level(l), ref.level(l))
    with pytest.raises(ValueError):
        ops.CorrOperands(1, 24, 40, 24, 40, 256, 4, 'tf32', cuda)


def test_avgpool_nhwc(cuda):
    from sd_animation_optical_flow_b200 import ops
    x = torch.randn((2, 9, 11, 32), device=cuda)
    ref = F.avg_pool2d(x.permute(0, 3, 1, 2), 2, stride=2).permute(0, 2, 3, 1)
    assert torch.allclose(ops.avgpool2_nhwc(x), ref, atol=1e-6)
