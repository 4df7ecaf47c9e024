"""ctypes binding of libsdof_b200.so (include/sdof_b200.h).

There is no fallback: if the library is missing or a call fails this module raises.
Torch tensors are only used as device-memory handles (`data_ptr()`), and the current
torch CUDA stream is passed to every call, so all work is stream-ordered with the
surrounding PyTorch ops.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int16, c_int32, c_int64, c_void_p

import torch

SDOF_MAX_LEVELS = 8
PRECISIONS = {'tf32': 0, '3xtf32': 1, 'bf16': 2, 'fp32': 3, 'fp16': 4}

_LIB_PATH = os.environ.get('SDOF_B200_LIB') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libsdof_b200.so')   # SDOF_B200_LIB: A/B runs against another build


class PyramidLayout(ctypes.Structure):
    _fields_ = [
        ('levels', c_int32),
        ('h', c_int32 * SDOF_MAX_LEVELS),
        ('w', c_int32 * SDOF_MAX_LEVELS),
        ('wp', c_int32 * SDOF_MAX_LEVELS),
        ('pitch', c_int64 * SDOF_MAX_LEVELS),
        ('offset', c_int64 * SDOF_MAX_LEVELS),
        ('total_floats', c_int64),
    ]


# name -> (restype, argtypes); must list every symbol include/sdof_b200.h declares
_P = c_void_p
SIGNATURES = {
    'sdof_abi_version': (c_int, []),
    'sdof_last_error': (c_char_p, []),
    'sdof_launch_count': (c_int64, []),
    'sdof_warp_tile_stats': (c_int, [POINTER(c_int64), c_int]),
    'sdof_fastdiv_u31': (ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_uint32]),
    'sdof_corr_pyramid_layout': (c_int, [c_int64, c_int, c_int, c_int, POINTER(PyramidLayout)]),
    'sdof_corr_volume_workspace_bytes': (c_int64, [c_int] * 8),
    'sdof_corr_volume_pyramid': (c_int, [_P, _P] + [c_int] * 8 + [_P, _P, c_int64, _P]),
    'sdof_corr_prepare_operands': (c_int, [_P, _P] + [c_int] * 9 + [_P, c_int64, _P]),
    'sdof_corr_pyramid_from_operands': (c_int, [c_int] * 8 + [_P, _P, c_int64, _P]),
    'sdof_corr_pyramid_layout_ex': (c_int, [c_int64, c_int, c_int, c_int, c_int, POINTER(PyramidLayout)]),
    'sdof_corr_src_operand_bytes': (c_int64, [c_int] * 4),
    'sdof_corr_tgt_operand_bytes': (c_int64, [c_int] * 5),
    'sdof_corr_prepare_src': (c_int, [_P] + [c_int] * 5 + [_P, c_int64, _P]),
    'sdof_corr_prepare_tgt': (c_int, [_P] + [c_int] * 6 + [_P, c_int64, _P]),
    'sdof_corr_prepare_both': (c_int, [_P, c_int, c_int, c_int, _P, c_int64, _P, c_int, c_int, c_int, c_int, _P, c_int64, c_int, c_int, _P]),
    'sdof_corr_pyramid_from_parts': (c_int, [_P, _P] + [c_int] * 10 + [_P, _P]),
    'sdof_corr_lookup_ex': (c_int, [_P, c_int, _P] + [c_int] * 7 + [_P, c_int, _P]),
    'sdof_corr_lookup': (c_int, [_P, _P] + [c_int] * 7 + [_P, _P]),
    'sdof_corr_lookup_nhwc': (c_int, [_P, _P] + [c_int] * 7 + [_P, _P]),
    'sdof_relu_scatter': (c_int, [_P, _P, _P, c_int64, c_int, _P, c_int, c_int, _P, c_int, c_int, c_int, _P]),
    'sdof_gru_rh': (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    'sdof_gru_update': (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    'sdof_corr_lookup_h': (c_int, [_P, c_int, _P] + [c_int] * 7 + [_P, c_int, _P]),
    'sdof_conv7x7_c2_relu_h': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    'sdof_corr_lookup_gather_h': (c_int, [_P, c_int, _P, _P, c_float, c_float, _P, _P] + [c_int] * 7 + [_P, c_int, _P]),
    'sdof_conv7x7_c2_relu_coords_h': (c_int, [_P, _P, c_float, c_float, _P, _P, _P, c_int, c_int, c_int, _P]),
    'sdof_flow_im2col7_h': (c_int, [_P, _P, c_float, c_float, _P, c_int, c_int, c_int, c_int, _P]),
    'sdof_motion_tail16_h': (c_int, [_P, _P, _P, _P, c_int64, _P, c_int, _P]),
    'sdof_gru_rh_h': (c_int, [_P, c_int, _P, _P, _P, c_int64, _P]),
    'sdof_gru_update_h': (c_int, [_P, _P, _P, _P, _P, _P, c_int, _P, c_int64, _P]),
    'sdof_flowhead2_taps_h': (c_int, [_P, _P, c_int64, _P, _P]),
    'sdof_flowhead2_gather_update': (c_int, [_P, c_float, c_float, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    'sdof_motion_tail16': (c_int, [_P, _P, _P, _P, c_int64, _P, c_int, _P]),
    'sdof_gru_zr_tc': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    'sdof_gru_q_tc': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P]),
    'sdof_flow_update': (c_int, [_P, c_float, c_float, _P, _P, _P, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    'sdof_convex_upsample': (c_int, [_P, _P, c_float, _P, c_int, c_int, c_int, _P, _P]),
    'sdof_instnorm_relu_nchw': (c_int, [_P, _P, c_int64, c_int64, c_float, c_int, _P]),
    'sdof_instnorm_stats_nhwc': (c_int, [_P, c_int, c_int64, c_int, _P, _P]),
    'sdof_instnorm_apply_nhwc': (c_int, [_P, _P, _P, _P, c_int, c_int64, c_int, c_float, c_int, _P]),
    'sdof_instnorm_stats_nhwc_h': (c_int, [_P, c_int, c_int64, c_int, _P, _P]),
    'sdof_instnorm_apply_nhwc_h': (c_int, [_P, _P, _P, _P, c_int, c_int64, c_int, c_float, c_int, _P]),
    'sdof_add_relu': (c_int, [_P, _P, _P, c_int64, _P]),
    'sdof_normalize_pad_u8_nhwc': (c_int, [_P] + [c_int] * 9 + [_P, _P]),
    'sdof_conv7x7_c2_relu': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    'sdof_flowhead2_update': (c_int, [_P, _P, c_float, c_float, _P, _P, _P, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    'sdof_alt_corr_forward': (c_int, [_P, _P, _P] + [c_int] * 8 + [_P, _P]),
    'sdof_alt_corr_level': (c_int, [_P, _P, _P] + [c_int] * 7 + [c_float, c_float, c_int, c_int, _P, _P]),
    'sdof_avgpool2_nhwc': (c_int, [_P] + [c_int] * 4 + [_P, _P]),
    'sdof_warp_cubic_u8': (c_int, [_P, _P] + [c_int] * 7 + [c_float, _P, _P]),
    'sdof_warp_cubic_f32': (c_int, [_P, _P] + [c_int] * 7 + [c_float, _P, _P]),
    'sdof_warp_bilinear_u8': (c_int, [_P, _P] + [c_int] * 7 + [c_float, _P, _P]),
    'sdof_warp_bilinear_f32': (c_int, [_P, _P] + [c_int] * 7 + [c_float, _P, _P]),
    'sdof_resize_cubic_f32': (c_int, [_P] + [c_int] * 6 + [_P, _P]),
    'sdof_cubic_table_i16': (c_int, [POINTER(c_int16)]),
    'sdof_ellipse_half_widths': (c_int, [c_int, POINTER(c_int32)]),
    'sdof_confidence_softmax': (c_int, [_P] + [c_int] * 4 + [_P, _P, _P]),
    'sdof_travel_distance': (c_int, [_P, _P] + [c_int] * 3 + [c_float, _P, _P]),
    'sdof_generate_mask': (c_int, [_P, _P] + [c_int] * 3 + [c_float, c_int, _P, _P]),
    'sdof_dilate_ellipse_u8': (c_int, [_P] + [c_int] * 5 + [_P, _P]),
    'sdof_expand_mask': (c_int, [_P, _P] + [c_int] * 4 + [_P, _P, _P]),
    'sdof_mix_propagated': (c_int, [_P, _P, _P] + [c_int] * 4 + [c_float, _P, _P]),
    'sdof_merge_select': (c_int, [_P, _P, _P] + [c_int] * 4 + [_P, _P]),
    'sdof_greedy_workspace_bytes': (c_int64, [c_int] * 3),
    'sdof_greedy_composite': (c_int, [_P, _P] + [c_int] * 3 + [c_float, _P, _P, _P, _P, _P]),
    'sdof_confidence_sums': (c_int, [_P, c_int, c_int64, _P, _P]),
    'sdof_detect_edges_workspace_bytes': (c_int64, [c_int, c_int]),
    'sdof_detect_edges': (c_int, [_P] + [c_int] * 5 + [_P, _P, c_int64, _P]),
    'sdof_abs_diff_sum_u8': (c_int, [_P, _P, c_int64, _P, _P]),
    'sdof_mask_blur_composite': (c_int, [_P, _P, _P] + [c_int] * 4 + [c_float, _P, _P, _P]),
    'sdof_resize_bicubic_workspace_bytes': (c_int64, [c_int] * 5),
    'sdof_box_blur_params': (c_int, [c_float, POINTER(c_int32)]),
    'sdof_resample_table': (c_int, [c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), c_int64]),
    'sdof_resize_bicubic_u8': (c_int, [_P] + [c_int] * 5 + [_P, _P, _P, c_int64, _P]),
    'sdof_warp_mask_composite': (c_int, [_P, _P, _P, _P] + [c_int] * 4 + [c_float, c_int, _P, _P, _P]),
}

_lib = None


class SdofError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load libsdof_b200.so and bind every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise SdofError(
            f'{_LIB_PATH} is missing: build it with `python -m sd_animation_optical_flow_b200.build` '
            '(needs nvcc with sm_100a support).  There is no CPU or PyTorch fallback for this path.')
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.sdof_abi_version() != 1:
        raise SdofError(f'libsdof_b200.so ABI {lib.sdof_abi_version()} != 1 expected by this package')
    _lib = lib
    return lib


_tls = threading.local()


def check(rc: int, what: str) -> None:
    """Status check of a library call; also switches back to the device that was current before `stream_ptr` made the
    tensors' device current for this call (see stream_ptr)."""
    prev = getattr(_tls, 'restore_device', None)
    if prev is not None:
        _tls.restore_device = None
        torch.cuda.set_device(prev)
    if rc != 0:
        msg = load().sdof_last_error()
        raise SdofError(f'{what} failed (status {rc}): {msg.decode() if msg else "?"}')


def stream_ptr(device=None) -> c_void_p:
    """Current torch stream of `device`, the stream handle every entry point takes.  The library launches on the CURRENT
    CUDA device and keys its per-device state (weight tables, SM count, kernel attributes) on it, so when the tensors'
    device is not the current one it is made current here, for the duration of the call: every wrapper is written as
    `check(lib.fn(..., stream_ptr(dev)), name)`, arguments are evaluated before the call, and `check` switches back.
    (The reference's PDCNetAux(device=cuda:N), ofgen_keyframe_inpaint.py:550,1128, runs on a non-default device.)"""
    if device is not None:
        idx = torch.device(device).index
        cur = torch.cuda.current_device()
        if idx is not None and idx != cur:
            if getattr(_tls, 'restore_device', None) is None:
                _tls.restore_device = cur
            torch.cuda.set_device(idx)
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t) -> c_void_p:
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def require_cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    """Mirror of the reference's CHECK_INPUT (RAFT/alt_cuda_corr/correlation.cpp:19-21),
    plus the dtype check the reference leaves to a hard accessor failure."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor')
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be contiguous')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f'{name} must have dtype {dtype}, got {t.dtype}')
    return t


def pyramid_layout(rows: int, h2: int, w2: int, levels: int, elem_bytes: int = 4) -> PyramidLayout:
    """Layout of a pyramid of fp32 (elem_bytes 4) or fp16 (2) elements; offsets / pitches / total_floats count ELEMENTS."""
    lay = PyramidLayout()
    check(load().sdof_corr_pyramid_layout_ex(rows, h2, w2, levels, elem_bytes, ctypes.byref(lay)), 'sdof_corr_pyramid_layout_ex')
    return lay


def launch_count() -> int:
    return int(load().sdof_launch_count())
