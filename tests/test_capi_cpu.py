"""The C-ABI library on a machine WITHOUT a GPU: it must load, export every symbol
include/sdof_b200.h declares, validate arguments, and its host-side tables must equal the
oracle's (and therefore OpenCV's).  No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import mask_oracle, warp_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, 'include', 'sdof_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sdof_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_header_symbol(lib):
    from sd_animation_optical_flow_b200 import _capi
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/sdof_b200.h but not exported'
    # and the Python binding covers exactly the header
    assert sorted(_capi.SIGNATURES) == names


def test_abi_version(lib):
    assert lib.sdof_abi_version() == 1


def test_cubic_table_matches_oracle(lib):
    from sd_animation_optical_flow_b200 import ops
    assert np.array_equal(ops.cubic_table_i16(), warp_oracle.cubic_table_i16())


def test_ellipse_rows_match_oracle(lib):
    from sd_animation_optical_flow_b200 import ops
    for k in (1, 3, 5, 7, 9, 15, 21, 31):
        assert ops.ellipse_half_widths(k) == mask_oracle.ellipse_half_widths(k)
    from sd_animation_optical_flow_b200._capi import SdofError
    with pytest.raises(SdofError):
        ops.ellipse_half_widths(8)


def test_pyramid_layout(lib):
    from sd_animation_optical_flow_b200 import _capi
    lay = _capi.pyramid_layout(6144, 96, 64, 4)
    assert list(lay.h[:4]) == [96, 48, 24, 12] and list(lay.w[:4]) == [64, 32, 16, 8]
    assert list(lay.wp[:4]) == [64, 32, 16, 8]
    # 200.5 MB for the 768x512 pyramid (+ < 2 % pitch padding)
    algo = 6144 * (96 * 64 + 48 * 32 + 24 * 16 + 12 * 8) * 4
    assert algo <= lay.total_floats * 4 <= algo * 1.02
    lay = _capi.pyramid_layout(14400, 90, 160, 4)   # 720x1280: floor pooling drops odd rows
    assert list(lay.h[:4]) == [90, 45, 22, 11] and list(lay.w[:4]) == [160, 80, 40, 20]
    lay = _capi.pyramid_layout(10, 18, 22, 4)       # rows padded to 4 floats for TMA
    assert list(lay.w[:4]) == [22, 11, 5, 2] and list(lay.wp[:4]) == [24, 12, 8, 4]
    for l in range(4):
        assert lay.offset[l] % 32 == 0 and lay.pitch[l] == lay.h[l] * lay.wp[l]
    with pytest.raises(_capi.SdofError):
        _capi.pyramid_layout(10, 18, 22, 9)


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected before any CUDA call, with a message."""
    null = ctypes.c_void_p(0)
    rc = lib.sdof_warp_cubic_u8(null, null, 1, 1, 8, 8, 3, 8, 8, 1.0, null, null)
    assert rc == 1 and b'NULL' in lib.sdof_last_error()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.sdof_warp_cubic_u8(p, p, 1, 1, 8, 8, 7, 8, 8, 1.0, p, null)
    assert rc == 1 and b'C must be in 1..4' in lib.sdof_last_error()
    rc = lib.sdof_warp_cubic_u8(p, p, 1, 1, 8, 8, 3, 8, 8, 0.5, p, null)
    assert rc == 1 and b'sign' in lib.sdof_last_error()
    rc = lib.sdof_corr_volume_pyramid(p, p, 1, 4, 4, 4, 4, 6, 4, 0, p, null, 0, null)
    assert rc == 1 and b'multiple of 4' in lib.sdof_last_error()
    rc = lib.sdof_corr_lookup(p, p, 1, 4, 4, 4, 4, 4, 99, p, null)
    assert rc == 1 and b'radius' in lib.sdof_last_error()
    rc = lib.sdof_dilate_ellipse_u8(p, 1, 8, 8, 4, 0, ctypes.cast((ctypes.c_float * 64)(), ctypes.c_void_p), null)
    assert rc == 1 and b'ksize' in lib.sdof_last_error()
    rc = lib.sdof_alt_corr_forward(p, p, p, 1, 4, 4, 4, 4, 6, 1, 4, p, null)
    assert rc == 1


def test_workspace_sizes(lib):
    n = 6144
    ws = lambda prec: lib.sdof_corr_volume_workspace_bytes(1, 96, 64, 96, 64, 256, 4, prec)
    assert ws(3) == 0                       # fp32: none
    assert ws(0) == 2 * n * 256 * 4         # tf32: rounded copies
    assert ws(1) == 4 * n * 256 * 4         # 3xtf32: hi + lo
    pooled = n + n // 4 + n // 16 + n // 64
    hdr = 2 * 2048                          # one scale header per operand buffer (source, target)
    assert ws(4) == max((n + pooled) * 256 * 2 + hdr, 2 * n * 256 * 4)   # fp16: 16-bit fmap1 + pooled fmap2 levels (or the tf32 fallback)
    assert ws(2) == (n + pooled) * 256 * 2 + hdr                          # bf16
    assert lib.sdof_corr_src_operand_bytes(1, 96, 64, 256) + lib.sdof_corr_tgt_operand_bytes(1, 96, 64, 256, 4) == ws(2)
    assert lib.sdof_corr_tgt_operand_bytes(1, 96, 64, 256, 4) == pooled * 256 * 2 + 2048   # a shared key-frame target: batch 1


def test_fp16_pyramid_layout(lib):
    """sdof_corr_pyramid_layout_ex: elem_bytes 4 == the fp32 layout; elem_bytes 2 = rows padded to 8 halves (an odd width
    always has a spare column for the half2 pair store), levels 128-byte aligned, sizes in elements."""
    import ctypes
    from sd_animation_optical_flow_b200 import _capi
    a, b = _capi.PyramidLayout(), _capi.PyramidLayout()
    assert lib.sdof_corr_pyramid_layout(6144, 96, 64, 4, ctypes.byref(a)) == 0
    assert lib.sdof_corr_pyramid_layout_ex(6144, 96, 64, 4, 4, ctypes.byref(b)) == 0
    assert bytes(a) == bytes(b)
    assert lib.sdof_corr_pyramid_layout_ex(6144, 96, 64, 4, 2, ctypes.byref(b)) == 0
    assert list(b.w[:4]) == [64, 32, 16, 8] and list(b.wp[:4]) == [64, 32, 16, 8]
    assert b.offset[0] == 64                                                              # 128-byte header (the factor)
    assert b.total_floats * 2 == 6144 * (96 * 64 + 48 * 32 + 24 * 16 + 12 * 8) * 2 + 128  # 100.3 MB: half of the fp32 pyramid
    assert lib.sdof_corr_pyramid_layout_ex(396, 18, 22, 4, 2, ctypes.byref(b)) == 0
    assert list(b.w[:4]) == [22, 11, 5, 2] and list(b.wp[:4]) == [24, 16, 8, 8]
    for l in range(4):
        assert b.wp[l] > b.w[l] or b.w[l] % 2 == 0
        assert (b.offset[l] * 2) % 128 == 0
    assert lib.sdof_corr_pyramid_layout_ex(1, 8, 8, 4, 3, ctypes.byref(b)) != 0


def test_product_path_has_no_cpu_fallback():
    """CUDA-only operators must refuse CPU tensors instead of silently computing elsewhere."""
    import torch
    from sd_animation_optical_flow_b200 import alt_cuda_corr, corr, ops
    f = torch.zeros(1, 8, 4, 4)
    with pytest.raises(RuntimeError):
        corr.CorrBlock(f, f)
    with pytest.raises(RuntimeError):
        corr.AlternateCorrBlock(f, f)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        alt_cuda_corr.forward(torch.zeros(1, 4, 4, 8), torch.zeros(1, 4, 4, 8), torch.zeros(1, 1, 4, 4, 2), 4)
    with pytest.raises(RuntimeError):
        ops.warp(torch.zeros(4, 4, 3, dtype=torch.uint8), torch.zeros(4, 4, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'sd_animation_optical_flow_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f'{fn} imports the oracle'


def test_blur_and_resample_host_tables_match_the_oracle(lib):
    """The host-side tables of csrc/blur.cu (Pillow's box-blur weights and resample coefficients) against the oracle's,
    which tests/test_oracle_blur.py pins against Pillow itself.  No GPU needed."""
    import ctypes

    import numpy as np

    from oracle import blur_oracle as bo
    out3 = (ctypes.c_int32 * 3)()
    for r in (0.0, 0.1, 0.5, 1, 2, 3.3, 4, 5.5, 8, 12, 30, 64):
        assert lib.sdof_box_blur_params(float(r), out3) == 0
        want = bo.box_weights(bo.gaussian_box_radius(r))
        assert tuple(np.uint32(v) for v in out3) == tuple(np.uint32(v) for v in want), (r, list(out3), want)
    for (n_in, n_out) in ((768, 96), (512, 64), (1280, 160), (720, 90), (100, 12), (60, 7), (16, 16), (40, 80), (9, 1)):
        bounds, coeffs = bo.resample_coeffs(n_in, n_out)
        kmax = 64 * max(1, -(-n_in // n_out)) + 8
        b = (ctypes.c_int32 * (2 * n_out))()
        c = (ctypes.c_int32 * (n_out * kmax))()
        ks = ctypes.c_int32()
        assert lib.sdof_resample_table(n_in, n_out, ctypes.byref(ks), b, c, n_out * kmax) == 0
        for i in range(n_out):
            assert b[2 * i] == bounds[i] and b[2 * i + 1] == len(coeffs[i]), (n_in, n_out, i)
            assert list(c[i * ks.value:i * ks.value + len(coeffs[i])]) == [int(v) for v in coeffs[i]], (n_in, n_out, i)


def test_tile_coordinate_division_constants(lib):
    """warp_tiled.cuh::wt_make_div: floor(n / d) by one multiply-high and a shift must be exact for every 0 <= n < 2^31."""
    import random
    rnd = random.Random(0)
    ds = list(range(1, 300)) + [2 ** k for k in range(1, 31)] + [2 ** k - 1 for k in range(2, 31)] + [2 ** k + 1 for k in range(1, 30)] + \
        [rnd.randrange(1, 2 ** 31) for _ in range(300)]
    for d in ds:
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1, 2 ** 31 - 2, (2 ** 31 - 1) // d * d, (2 ** 31 - 1) // d * d - 1]
        ns += [rnd.randrange(0, 2 ** 31) for _ in range(40)]
        for n in ns:
            if 0 <= n < 2 ** 31:
                assert lib.sdof_fastdiv_u31(n, d) == n // d, (n, d)


def test_header_and_ctypes_binding_agree_on_every_signature():
    """Parameter count and pointer/scalar kind of every declaration in include/sdof_b200.h against _capi.SIGNATURES
    (an ABI drift between the documented header and the binding the product uses must fail here, without a GPU)."""
    from ctypes import c_void_p
    from sd_animation_optical_flow_b200 import _capi
    text = open(os.path.join(ROOT, 'include', 'sdof_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = re.findall(r'\b[a-z_0-9]+\s*\*?\s*(sdof_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S)
    assert len(decls) == len(_capi.SIGNATURES)
    for name, params in decls:
        params = ' '.join(params.split())
        plist = [] if params in ('', 'void') else [p.strip() for p in params.split(',')]
        _, argtypes = _capi.SIGNATURES[name]
        assert len(plist) == len(argtypes), f'{name}: header has {len(plist)} parameters, binding {len(argtypes)}'
        for p, a in zip(plist, argtypes):
            is_ptr_h = '*' in p or 'sdof_stream_t' in p
            is_ptr_b = a is c_void_p or hasattr(a, 'contents') or a is ctypes.c_char_p
            assert is_ptr_h == is_ptr_b, f'{name}: parameter "{p}" vs binding {a}'
