"""Pins oracle/corr_oracle.py against the reference's own CorrBlock (tests/golden/corr.npz, made by
importing RAFT/core/corr.py) and checks the on-the-fly (alt_cuda_corr) restatement against it."""
import numpy as np
import pytest

from oracle import corr_oracle as co
from tests import golden_inputs as gi


@pytest.mark.parametrize('name', list(gi.CORR_CASES))
def test_pyramid_and_lookup_match_reference(golden, name):
    f1, f2, coords = gi.corr_inputs(name)
    pyr = co.corr_pyramid(f1, f2, 4)
    for l, lv in enumerate(pyr):
        ref = golden['corr'][f'{name}_pyr{l}']
        rows = gi.PYRAMID_ROWS(lv.shape[0])
        assert lv[rows].shape == ref.shape
        np.testing.assert_allclose(lv[rows], ref, rtol=1e-5, atol=2e-5)
    look = co.corr_lookup(pyr, coords, 4)
    ref = golden['corr'][f'{name}_lookup']
    assert look.shape == ref.shape
    # 1e-5 abs on |corr| <= ~6
    np.testing.assert_allclose(look, ref, rtol=0, atol=3e-5)


@pytest.mark.parametrize('name', list(gi.CORR_CASES))
def test_alternate_corr_equals_corrblock(golden, name):
    f1, f2, coords = gi.corr_inputs(name)
    alt = co.alternate_corr_block(f1, f2, coords, 4, 4)
    np.testing.assert_allclose(alt, golden['corr'][f'{name}_lookup'], rtol=0, atol=5e-5)


def test_channel_order_is_x_major():
    """Looking up at integer coords: channel 9*ix+iy of level 0 is volume[p, y+iy-4, x+ix-4]."""
    f1, f2, _ = gi.corr_inputs('even')
    B, C, h, w = f1.shape
    pyr = co.corr_pyramid(f1, f2, 1)
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing='ij')
    coords = np.stack([xs, ys], 0).astype(np.float32)[None]
    look = co.corr_lookup(pyr, coords, 4)
    p = 5 * w + 7
    vol = pyr[0][p]
    for ix, iy in ((0, 0), (8, 0), (3, 6), (4, 4)):
        yy, xx = 5 + iy - 4, 7 + ix - 4
        want = vol[yy, xx] if 0 <= yy < h and 0 <= xx < w else 0.0
        assert look[0, 9 * ix + iy, 5, 7] == pytest.approx(want, abs=1e-6)
