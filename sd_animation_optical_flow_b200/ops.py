"""Device-resident operators over CUDA tensors: thin, checked wrappers of the C ABI.

These are the batched, stream-ordered, host-sync-free forms SURVEY §8(b) asks for next to
the reference-named numpy entry points (pdcnet_of.py / ofgen.py in this package):
`warp`, `confidence_softmax`, `generate_mask`, `composite`, ...  Every function launches
this package's sm_100a kernels through libsdof_b200.so and raises if that is impossible.
"""
from __future__ import annotations

import ctypes

import torch

from . import _capi
from ._capi import PRECISIONS, check, load, ptr, require_cuda, stream_ptr

f32, u8 = torch.float32, torch.uint8


# --------------------------------------------------------------------------- correlation
class CorrPyramid:
    """The correlation volume and its pooled levels in one HBM buffer (see sdof_pyramid_layout /
    sdof_corr_pyramid_layout_ex in include/sdof_b200.h): fp32, or fp16 (`buf.dtype == torch.float16`)."""

    def __init__(self, buf: torch.Tensor, layout, B: int, h1: int, w1: int, h2: int, w2: int, levels: int):
        self.buf, self.layout = buf, layout
        self.B, self.h1, self.w1, self.h2, self.w2, self.levels = B, h1, w1, h2, w2, levels

    @property
    def elem_bytes(self) -> int:
        return self.buf.element_size()

    def level(self, l: int) -> torch.Tensor:
        """View of level l shaped like the reference's corr_pyramid[l]:
        [B*h1*w1, 1, h_l, w_l] (RAFT/core/corr.py:21-27); strided when the row pitch is padded.
        For an fp16 pyramid these are the STORED values (raw accumulators of the auto-ranged operands): multiply by
        `factor`, or use `level_values`, for correlation units."""
        lay = self.layout
        rows = self.B * self.h1 * self.w1
        return torch.as_strided(self.buf, (rows, 1, lay.h[l], lay.w[l]), (lay.pitch[l], 0, lay.wp[l], 1), lay.offset[l])

    @property
    def factor(self) -> torch.Tensor:
        """0-d fp32 tensor: stored value * factor = correlation (1 for an fp32 pyramid; an fp16 pyramid keeps it in the
        128-byte header of the buffer, where the lookup kernel reads it)."""
        if self.buf.dtype == f32:
            return torch.ones((), dtype=f32, device=self.buf.device)
        return self.buf[:2].view(f32)[0]

    def level_values(self, l: int) -> torch.Tensor:
        """Level l in correlation units, fp32 (a copy for an fp16 pyramid)."""
        v = self.level(l)
        return v if v.dtype == f32 else v.float() * self.factor


STORAGE = {'fp32': (f32, 4), 'fp16': (torch.float16, 2)}


def _new_pyramid(B, h1, w1, h2, w2, levels, storage, device) -> CorrPyramid:
    dt, eb = STORAGE[storage]
    lay = _capi.pyramid_layout(B * h1 * w1, h2, w2, levels, eb)
    buf = torch.empty(max(int(lay.total_floats), 64), dtype=dt, device=device)
    return CorrPyramid(buf, lay, B, h1, w1, h2, w2, levels)


class CorrTarget:
    """Prepared 16-bit TARGET operand of the fp16/bf16 correlation path: fmap2 and its avg-pooled levels, auto-ranged
    (sdof_corr_prepare_tgt).  In the key-frame scheme fmap2 belongs to the key frame: build it once (batch 1) and hand it to
    every `corr_volume_pyramid(..., target=...)` / `CorrSource.pyramid` call of the pairs that share the key."""

    def __init__(self, fmap2_nhwc: torch.Tensor, levels: int = 4, precision: str = 'fp16', _prepare: bool = True):
        require_cuda(fmap2_nhwc, 'fmap2', f32)
        if precision not in ('fp16', 'bf16'):
            raise ValueError("prepared operands exist only for precision 'fp16' / 'bf16'")
        if fmap2_nhwc.dim() != 4:
            raise RuntimeError('fmap2 must be [B2,h2,w2,C]')
        self.B2, self.h2, self.w2, self.C = fmap2_nhwc.shape
        self.levels, self.precision = levels, precision
        lib = load()
        n = int(lib.sdof_corr_tgt_operand_bytes(self.B2, self.h2, self.w2, self.C, levels))
        self.buf = torch.empty(n, dtype=u8, device=fmap2_nhwc.device)
        if _prepare:
            check(lib.sdof_corr_prepare_tgt(ptr(fmap2_nhwc), self.B2, self.h2, self.w2, self.C, levels, PRECISIONS[precision],
                                            ptr(self.buf), n, stream_ptr(self.buf.device)), 'sdof_corr_prepare_tgt')


class CorrSource:
    """Prepared 16-bit SOURCE operand (fmap1, auto-ranged; sdof_corr_prepare_src)."""

    def __init__(self, fmap1_nhwc: torch.Tensor, precision: str = 'fp16', _prepare: bool = True):
        require_cuda(fmap1_nhwc, 'fmap1', f32)
        if precision not in ('fp16', 'bf16'):
            raise ValueError("prepared operands exist only for precision 'fp16' / 'bf16'")
        if fmap1_nhwc.dim() != 4:
            raise RuntimeError('fmap1 must be [B,h1,w1,C]')
        self.B, self.h1, self.w1, self.C = fmap1_nhwc.shape
        self.precision = precision
        lib = load()
        n = int(lib.sdof_corr_src_operand_bytes(self.B, self.h1, self.w1, self.C))
        self.buf = torch.empty(n, dtype=u8, device=fmap1_nhwc.device)
        if _prepare:
            check(lib.sdof_corr_prepare_src(ptr(fmap1_nhwc), self.B, self.h1, self.w1, self.C, PRECISIONS[precision], ptr(self.buf), n,
                                            stream_ptr(self.buf.device)), 'sdof_corr_prepare_src')

    def pyramid(self, target: CorrTarget, storage: str = 'fp16', out: CorrPyramid | None = None) -> CorrPyramid:
        """The tcgen05 kernel alone on prepared operands; target.B2 must be B or 1 (shared key frame)."""
        if target.C != self.C or target.precision != self.precision or target.B2 not in (1, self.B):
            raise RuntimeError('source and target operands disagree (channels / precision / batch)')
        if out is None:
            out = _new_pyramid(self.B, self.h1, self.w1, target.h2, target.w2, target.levels, storage, self.buf.device)
        check(load().sdof_corr_pyramid_from_parts(ptr(self.buf), ptr(target.buf), self.B, self.h1, self.w1, target.B2, target.h2,
                                                  target.w2, self.C, target.levels, PRECISIONS[self.precision], out.elem_bytes,
                                                  ptr(out.buf), stream_ptr(self.buf.device)), 'sdof_corr_pyramid_from_parts')
        return out


def prepare_pair(fmap1_nhwc: torch.Tensor, fmap2_nhwc: torch.Tensor, levels: int = 4, precision: str = 'fp16'):
    """(CorrSource, CorrTarget) of one pair (or B pairs, or B sources against one target) prepared by ONE abs-max launch and
    ONE conversion launch (sdof_corr_prepare_both) instead of two each."""
    target = CorrTarget(fmap2_nhwc, levels, precision, _prepare=False)
    source = CorrSource(fmap1_nhwc, precision, _prepare=False)
    if source.C != target.C:
        raise RuntimeError('fmap1 and fmap2 disagree on channels')
    if source.B and target.B2:
        check(load().sdof_corr_prepare_both(ptr(fmap1_nhwc), source.B, source.h1, source.w1, ptr(source.buf), source.buf.numel(),
                                            ptr(fmap2_nhwc), target.B2, target.h2, target.w2, levels, ptr(target.buf),
                                            target.buf.numel(), source.C, PRECISIONS[precision], stream_ptr(fmap1_nhwc.device)),
              'sdof_corr_prepare_both')
    return source, target


def corr_volume_pyramid(fmap1_nhwc: torch.Tensor, fmap2_nhwc: torch.Tensor | None, levels: int = 4,
                        precision: str = 'fp16', storage: str = 'fp32', target: CorrTarget | None = None) -> CorrPyramid:
    """fmap1 [B,h1,w1,C], fmap2 [B,h2,w2,C] fp32 channels-last -> CorrPyramid (C1 + C2).
    storage 'fp16' (fp16 / bf16 precision only) stores the pyramid in half precision; `target` = a prepared CorrTarget
    (then fmap2 is ignored), which may have batch 1 = one key frame shared by all B pairs."""
    require_cuda(fmap1_nhwc, 'fmap1', f32)
    if fmap1_nhwc.dim() != 4:
        raise RuntimeError('feature maps must be [B,h,w,C]')
    if precision not in PRECISIONS:
        raise ValueError(f'precision must be one of {sorted(PRECISIONS)}, got {precision!r}')
    if storage not in STORAGE:
        raise ValueError(f"storage must be 'fp32' or 'fp16', got {storage!r}")
    B, h1, w1, C = fmap1_nhwc.shape
    if target is not None or storage == 'fp16':
        if precision not in ('fp16', 'bf16'):
            raise ValueError("fp16 pyramid storage / prepared targets need precision 'fp16' or 'bf16'")
        if target is None:
            require_cuda(fmap2_nhwc, 'fmap2', f32)
            if fmap2_nhwc.dim() != 4 or fmap2_nhwc.shape[0] not in (1, B) or fmap2_nhwc.shape[3] != C:
                raise RuntimeError(f'fmap1 {tuple(fmap1_nhwc.shape)} and fmap2 {tuple(fmap2_nhwc.shape)} disagree on batch/channels')
            if B == 0:
                return _new_pyramid(B, h1, w1, fmap2_nhwc.shape[1], fmap2_nhwc.shape[2], levels, storage, fmap1_nhwc.device)
            source, target = prepare_pair(fmap1_nhwc, fmap2_nhwc, levels, precision)
            return source.pyramid(target, storage)
        if B == 0:
            return _new_pyramid(B, h1, w1, target.h2, target.w2, target.levels, storage, fmap1_nhwc.device)
        return CorrSource(fmap1_nhwc, precision).pyramid(target, storage)
    require_cuda(fmap2_nhwc, 'fmap2', f32)
    if fmap2_nhwc.dim() != 4:
        raise RuntimeError('feature maps must be [B,h,w,C]')
    B2, h2, w2, C2 = fmap2_nhwc.shape
    if B != B2 or C != C2:
        raise RuntimeError(f'fmap1 {tuple(fmap1_nhwc.shape)} and fmap2 {tuple(fmap2_nhwc.shape)} disagree on batch/channels')
    lib = load()
    dev = fmap1_nhwc.device
    pyr = _new_pyramid(B, h1, w1, h2, w2, levels, 'fp32', dev)
    ws_bytes = int(lib.sdof_corr_volume_workspace_bytes(B, h1, w1, h2, w2, C, levels, PRECISIONS[precision]))
    ws = torch.empty(ws_bytes, dtype=u8, device=dev) if ws_bytes else None
    check(lib.sdof_corr_volume_pyramid(ptr(fmap1_nhwc), ptr(fmap2_nhwc), B, h1, w1, h2, w2, C, levels,
                                       PRECISIONS[precision], ptr(pyr.buf), ptr(ws), ws_bytes, stream_ptr(dev)),
          'sdof_corr_volume_pyramid')
    return pyr


class CorrOperands:
    """Round-1 single-workspace form of the prepared operands (sdof_corr_prepare_operands / sdof_corr_pyramid_from_operands,
    fp32 pyramid).  New code uses CorrSource / CorrTarget."""

    def __init__(self, B, h1, w1, h2, w2, C, levels, precision, device):
        if precision not in ('fp16', 'bf16'):
            raise ValueError("prepared operands exist only for precision 'fp16' / 'bf16'")
        self.shape = (B, h1, w1, h2, w2, C)
        self.levels, self.precision = levels, precision
        n = int(load().sdof_corr_volume_workspace_bytes(B, h1, w1, h2, w2, C, levels, PRECISIONS[precision]))
        self.workspace = torch.empty(n, dtype=u8, device=device)

    def prepare(self, fmap1_nhwc: torch.Tensor | None = None, fmap2_nhwc: torch.Tensor | None = None):
        B, h1, w1, h2, w2, C = self.shape
        parts = 0
        if fmap1_nhwc is not None:
            require_cuda(fmap1_nhwc, 'fmap1', f32)
            assert tuple(fmap1_nhwc.shape) == (B, h1, w1, C)
            parts |= 1
        if fmap2_nhwc is not None:
            require_cuda(fmap2_nhwc, 'fmap2', f32)
            assert tuple(fmap2_nhwc.shape) == (B, h2, w2, C)
            parts |= 2
        if parts:
            check(load().sdof_corr_prepare_operands(ptr(fmap1_nhwc), ptr(fmap2_nhwc), B, h1, w1, h2, w2, C, self.levels,
                                                    PRECISIONS[self.precision], parts, ptr(self.workspace),
                                                    self.workspace.numel(), stream_ptr(self.workspace.device)),
                  'sdof_corr_prepare_operands')
        return self

    def pyramid(self, out: 'CorrPyramid | None' = None) -> 'CorrPyramid':
        B, h1, w1, h2, w2, C = self.shape
        if out is None:
            out = _new_pyramid(B, h1, w1, h2, w2, self.levels, 'fp32', self.workspace.device)
        check(load().sdof_corr_pyramid_from_operands(B, h1, w1, h2, w2, C, self.levels, PRECISIONS[self.precision], ptr(out.buf),
                                                     ptr(self.workspace), self.workspace.numel(),
                                                     stream_ptr(self.workspace.device)), 'sdof_corr_pyramid_from_operands')
        return out


def corr_lookup(pyr: CorrPyramid, coords: torch.Tensor, radius: int = 4, out: torch.Tensor | None = None) -> torch.Tensor:
    """coords [B,2,h1,w1] -> [B, levels*(2r+1)^2, h1, w1] (C3); fp32 output from an fp32 or fp16 pyramid."""
    require_cuda(coords, 'coords', f32)
    if tuple(coords.shape) != (pyr.B, 2, pyr.h1, pyr.w1):
        raise RuntimeError(f'coords must be {(pyr.B, 2, pyr.h1, pyr.w1)}, got {tuple(coords.shape)}')
    ch = pyr.levels * (2 * radius + 1) ** 2
    if out is None:
        out = torch.empty((pyr.B, ch, pyr.h1, pyr.w1), dtype=f32, device=coords.device)
    else:
        require_cuda(out, 'out', f32)
    if pyr.B == 0:
        return out
    check(load().sdof_corr_lookup_ex(ptr(pyr.buf), pyr.elem_bytes, ptr(coords), pyr.B, pyr.h1, pyr.w1, pyr.h2, pyr.w2, pyr.levels,
                                     radius, ptr(out), 0, stream_ptr(coords.device)), 'sdof_corr_lookup_ex')
    return out


def alt_corr_forward(fmap1: torch.Tensor, fmap2: torch.Tensor, coords: torch.Tensor, radius: int) -> torch.Tensor:
    """K1 with the reference op's contract: fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C],
    coords [B,N,H1,W1,2] -> [B,N,(2r+1)^2,H1,W1], unnormalised."""
    require_cuda(fmap1, 'fmap1', f32)
    require_cuda(fmap2, 'fmap2', f32)
    require_cuda(coords, 'coords', f32)
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    N = coords.shape[1]
    if tuple(coords.shape) != (B, N, H1, W1, 2):
        raise RuntimeError(f'coords must be [B,N,H1,W1,2], got {tuple(coords.shape)}')
    out = torch.empty((B, N, (2 * radius + 1) ** 2, H1, W1), dtype=f32, device=fmap1.device)
    check(load().sdof_alt_corr_forward(ptr(fmap1), ptr(fmap2), ptr(coords), B, H1, W1, H2, W2, C, N, radius, ptr(out),
                                       stream_ptr(fmap1.device)), 'sdof_alt_corr_forward')
    return out


def alt_corr_level(fmap1: torch.Tensor, fmap2_level: torch.Tensor, coords: torch.Tensor, radius: int, coord_scale: float,
                   out_scale: float, chan_offset: int, out: torch.Tensor) -> torch.Tensor:
    """One level of C4 written into channels [chan_offset, +(2r+1)^2) of out [B,CH,H1,W1]."""
    require_cuda(fmap1, 'fmap1', f32)
    require_cuda(fmap2_level, 'fmap2', f32)
    require_cuda(coords, 'coords', f32)
    require_cuda(out, 'out', f32)
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2_level.shape
    check(load().sdof_alt_corr_level(ptr(fmap1), ptr(fmap2_level), ptr(coords), B, H1, W1, H2, W2, C, radius,
                                     float(coord_scale), float(out_scale), chan_offset, out.shape[1], ptr(out),
                                     stream_ptr(fmap1.device)), 'sdof_alt_corr_level')
    return out


def avgpool2_nhwc(x: torch.Tensor) -> torch.Tensor:
    require_cuda(x, 'x', f32)
    B, H, W, C = x.shape
    out = torch.empty((B, H // 2, W // 2, C), dtype=f32, device=x.device)
    check(load().sdof_avgpool2_nhwc(ptr(x), B, H, W, C, ptr(out), stream_ptr(x.device)), 'sdof_avgpool2_nhwc')
    return out


# --------------------------------------------------------------------------- warp
def warp(src: torch.Tensor, flow: torch.Tensor, mode: str = 'cv2_cubic', sign: float = 1.0) -> torch.Tensor:
    """Backward warp on the device.  src [Bs,Hs,Ws,C] (Bs == B or 1) or [Hs,Ws,C] / [Hs,Ws],
    u8 or f32; flow [B,H,W,2] or [H,W,2] f32.  mode 'cv2_cubic' (bit-exact cv2.remap for u8)
    or 'bilinear' (grid_sample semantics).  sign=+1 samples at x+flow (pdcnet_of.warp_frame),
    -1 at x-flow (ofgen.warp_frame)."""
    require_cuda(src, 'src')
    require_cuda(flow, 'flow', f32)
    squeeze_b = flow.dim() == 3
    fl = flow[None] if squeeze_b else flow
    s = src
    squeeze_c = False
    if s.dim() == 2:
        s, squeeze_c = s[None, :, :, None], True
    elif s.dim() == 3:
        if squeeze_b:
            s = s[None]
        else:  # [B,H,W] batch of single-channel images
            s, squeeze_c = s[..., None], True
    if s.dim() != 4 or fl.dim() != 4 or fl.shape[-1] != 2:
        raise RuntimeError(f'bad shapes: src {tuple(src.shape)}, flow {tuple(flow.shape)}')
    B, H, W, _ = fl.shape
    Bs, Hs, Ws, C = s.shape
    if Bs not in (1, B):
        raise RuntimeError(f'src batch {Bs} must be 1 or {B}')
    if mode not in ('cv2_cubic', 'bilinear'):
        raise ValueError(f"mode must be 'cv2_cubic' or 'bilinear', got {mode!r}")
    if s.dtype not in (u8, f32):
        raise RuntimeError(f'src must be uint8 or float32, got {s.dtype}')
    s = s.contiguous()
    out = torch.empty((B, H, W, C), dtype=s.dtype, device=s.device)
    if B == 0:   # empty batch: nothing to launch (empty tensors have no device pointer)
        return out[..., 0] if squeeze_c else out
    lib = load()
    fn = {('cv2_cubic', u8): lib.sdof_warp_cubic_u8, ('cv2_cubic', f32): lib.sdof_warp_cubic_f32,
          ('bilinear', u8): lib.sdof_warp_bilinear_u8, ('bilinear', f32): lib.sdof_warp_bilinear_f32}[(mode, s.dtype)]
    check(fn(ptr(s), ptr(fl), B, int(Bs == B and B > 0), Hs, Ws, C, H, W, float(sign), ptr(out), stream_ptr(s.device)),
          f'sdof_warp_{mode}')
    if squeeze_c:
        out = out[..., 0]
    return out[0] if squeeze_b else out


def resize_cubic(src: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """cv2.resize(src, (ow, oh), interpolation=cv2.INTER_CUBIC) of float32 channels-last images [B,H,W,C] / [H,W,C] / [H,W]
    on the device (W3: the latent resizes of warp_frame_latent, pdcnet_of.py:24,30)."""
    require_cuda(src, 'src', f32)
    s = src
    if s.dim() == 2:
        s = s[None, :, :, None]
    elif s.dim() == 3:
        s = s[None]
    if s.dim() != 4:
        raise RuntimeError(f'src must be [B,H,W,C], [H,W,C] or [H,W], got {tuple(src.shape)}')
    B, H, W, C = s.shape
    if oh < 1 or ow < 1:
        raise RuntimeError('resize_cubic: bad output size')
    out = torch.empty((B, oh, ow, C), dtype=f32, device=s.device)
    if B:
        check(load().sdof_resize_cubic_f32(ptr(s), B, H, W, C, oh, ow, ptr(out), stream_ptr(s.device)), 'sdof_resize_cubic_f32')
    if src.dim() == 2:
        return out[0, :, :, 0]
    return out[0] if src.dim() == 3 else out


def warp_tile_stats(reset: bool = True):
    """(staged, fallback) tile counts of the tiled u8 cubic warp kernel on the current device since the last reset
    (diagnostic; synchronises the device)."""
    buf = (ctypes.c_int64 * 2)()
    check(load().sdof_warp_tile_stats(buf, int(reset)), 'sdof_warp_tile_stats')
    return int(buf[0]), int(buf[1])


# --------------------------------------------------------------------------- confidence / masks
def confidence_softmax(weight_map: torch.Tensor):
    """weight_map [B,K,H,W] -> (confidence, log_confidence) [B,H,W] (M1)."""
    require_cuda(weight_map, 'weight_map', f32)
    B, K, H, W = weight_map.shape
    conf = torch.empty((B, H, W), dtype=f32, device=weight_map.device)
    logc = torch.empty_like(conf)
    check(load().sdof_confidence_softmax(ptr(weight_map), B, K, H, W, ptr(conf), ptr(logc), stream_ptr(weight_map.device)),
          'sdof_confidence_softmax')
    return conf, logc


def travel_distance(flow: torch.Tensor, conf: torch.Tensor, conf_thres: float = 0.9) -> torch.Tensor:
    require_cuda(flow, 'flow', f32)
    require_cuda(conf, 'conf', f32)
    B, H, W, _ = flow.shape
    v = torch.empty((B, H, W), dtype=f32, device=flow.device)
    check(load().sdof_travel_distance(ptr(flow), ptr(conf), B, H, W, float(conf_thres), ptr(v), stream_ptr(flow.device)),
          'sdof_travel_distance')
    return v


def generate_mask(conf: torch.Tensor, log_conf: torch.Tensor | None, thres: float, ksize: int = 7) -> torch.Tensor:
    """conf [B,H,W] -> dilated u8 mask; log_conf (if given) is reset in place where conf < thres (M3)."""
    require_cuda(conf, 'conf', f32)
    if log_conf is not None:
        require_cuda(log_conf, 'log_conf', f32)
    B, H, W = conf.shape
    mask = torch.empty((B, H, W), dtype=u8, device=conf.device)
    check(load().sdof_generate_mask(ptr(conf), ptr(log_conf), B, H, W, float(thres), ksize, ptr(mask), stream_ptr(conf.device)),
          'sdof_generate_mask')
    return mask


def dilate_ellipse(mask: torch.Tensor, ksize: int = 7, invert: bool = False) -> torch.Tensor:
    require_cuda(mask, 'mask', u8)
    B, H, W = mask.shape
    out = torch.empty_like(mask)
    check(load().sdof_dilate_ellipse_u8(ptr(mask), B, H, W, ksize, int(invert), ptr(out), stream_ptr(mask.device)),
          'sdof_dilate_ellipse_u8')
    return out


def expand_mask(mask: torch.Tensor, image: torch.Tensor, ksize: int = 7) -> torch.Tensor:
    require_cuda(mask, 'mask', u8)
    require_cuda(image, 'image', u8)
    B, H, W = mask.shape
    if tuple(image.shape) != (B, H, W, 3):
        raise RuntimeError(f'image must be {(B, H, W, 3)}, got {tuple(image.shape)}')
    scratch = torch.empty_like(mask)
    out = torch.empty_like(mask)
    check(load().sdof_expand_mask(ptr(mask), ptr(image), B, H, W, ksize, ptr(scratch), ptr(out), stream_ptr(mask.device)),
          'sdof_expand_mask')
    return out


def mix_propagated(raw: torch.Tensor, warped: torch.Tensor, mask: torch.Tensor, ppw: float) -> torch.Tensor:
    require_cuda(raw, 'raw', u8)
    require_cuda(warped, 'warped', u8)
    require_cuda(mask, 'mask', u8)
    B, H, W, C = raw.shape
    out = torch.empty_like(raw)
    check(load().sdof_mix_propagated(ptr(raw), ptr(warped), ptr(mask), B, H, W, C, float(ppw), ptr(out), stream_ptr(raw.device)),
          'sdof_mix_propagated')
    return out


def merge_select(base: torch.Tensor, second: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    require_cuda(base, 'base', u8)
    require_cuda(second, 'second', u8)
    require_cuda(mask, 'mask', u8)
    B, H, W, C = base.shape
    out = torch.empty_like(base)
    check(load().sdof_merge_select(ptr(base), ptr(second), ptr(mask), B, H, W, C, ptr(out), stream_ptr(base.device)),
          'sdof_merge_select')
    return out


def greedy_composite(flow_mat: torch.Tensor, ai_frames: torch.Tensor, thres: float):
    """flow_mat [n,H,W,3] f32 (modified in place like the reference), ai_frames [n,H,W,3] u8 ->
    (ret [H,W,3] u8, mask [H,W] u8, order [n] int32) (M5)."""
    require_cuda(flow_mat, 'flow_mat', f32)
    require_cuda(ai_frames, 'ai_frames', u8)
    n, H, W, three = flow_mat.shape
    if three != 3 or tuple(ai_frames.shape) != (n, H, W, 3):
        raise RuntimeError('flow_mat must be [n,H,W,3] and ai_frames [n,H,W,3]')
    dev = flow_mat.device
    ret = torch.empty((H, W, 3), dtype=u8, device=dev)
    mask = torch.empty((H, W), dtype=u8, device=dev)
    order = torch.empty((n,), dtype=torch.int32, device=dev)
    lib = load()
    ws = torch.empty(int(lib.sdof_greedy_workspace_bytes(n, H, W)), dtype=u8, device=dev)
    check(lib.sdof_greedy_composite(ptr(flow_mat), ptr(ai_frames), n, H, W, float(thres), ptr(ret), ptr(mask), ptr(order),
                                    ptr(ws), stream_ptr(dev)), 'sdof_greedy_composite')
    return ret, mask, order


def confidence_sums(flow_mat: torch.Tensor) -> torch.Tensor:
    """flow_mat [S, ..., 3] -> float64 [S] sums of the confidence channel (A2)."""
    require_cuda(flow_mat, 'flow_mat', f32)
    S = flow_mat.shape[0]
    per = flow_mat[0].numel() // 3 if S else 0
    sums = torch.empty((S,), dtype=torch.float64, device=flow_mat.device)
    check(load().sdof_confidence_sums(ptr(flow_mat), S, per, ptr(sums), stream_ptr(flow_mat.device)), 'sdof_confidence_sums')
    return sums


def warp_mask_composite(src: torch.Tensor, base: torch.Tensor, flow: torch.Tensor, weight_map: torch.Tensor,
                        thres: float, ksize: int = 7):
    """Fused W1+M1+M3+M4(ppw=1): src [Bs,H,W,3] u8 key frame(s), base [B,H,W,3] u8, flow [B,H,W,2],
    weight_map [B,2,H,W] -> (out [B,H,W,3] u8, mask [B,H,W] u8)."""
    require_cuda(src, 'src', u8)
    require_cuda(base, 'base', u8)
    require_cuda(flow, 'flow', f32)
    require_cuda(weight_map, 'weight_map', f32)
    B, H, W, _ = flow.shape
    if weight_map.shape[1] != 2:
        raise RuntimeError('weight_map must have 2 mixture components')
    if src.shape[0] not in (1, B) or tuple(src.shape[1:]) != (H, W, 3) or tuple(base.shape) != (B, H, W, 3):
        raise RuntimeError('src must be [B or 1,H,W,3] and base [B,H,W,3]')
    out = torch.empty_like(base)
    mask = torch.empty((B, H, W), dtype=u8, device=base.device)
    if B == 0:
        return out, mask
    check(load().sdof_warp_mask_composite(ptr(src), ptr(base), ptr(flow), ptr(weight_map), B, int(src.shape[0] == B), H, W,
                                          float(thres), ksize, ptr(out), ptr(mask), stream_ptr(base.device)),
          'sdof_warp_mask_composite')
    return out, mask


def cubic_table_i16():
    """Host copy of the kernel's bicubic weight table (no GPU needed)."""
    import numpy as np
    buf = (ctypes.c_int16 * (1024 * 16))()
    check(load().sdof_cubic_table_i16(buf), 'sdof_cubic_table_i16')
    return np.frombuffer(buf, dtype=np.int16).reshape(1024, 16).copy()


def ellipse_half_widths(ksize: int):
    buf = (ctypes.c_int32 * ksize)()
    check(load().sdof_ellipse_half_widths(ksize, buf), 'sdof_ellipse_half_widths')
    return list(buf)


# --------------------------------------------------------------------------- RAFT update-loop glue (NHWC)
def corr_lookup_nhwc(pyr: CorrPyramid, coords_nhwc: torch.Tensor, radius: int, out: torch.Tensor) -> torch.Tensor:
    """coords [B,h1,w1,2] -> out [B,h1,w1,levels*(2r+1)^2], channels-last."""
    check(load().sdof_corr_lookup_ex(ptr(pyr.buf), pyr.elem_bytes, ptr(coords_nhwc), pyr.B, pyr.h1, pyr.w1, pyr.h2, pyr.w2,
                                     pyr.levels, radius, ptr(out), 1, stream_ptr(out.device)), 'sdof_corr_lookup_ex')
    return out


def relu_scatter(src: torch.Tensor, dst1: torch.Tensor, off1: int, dst2: torch.Tensor | None = None, off2: int = 0,
                 c_valid: int | None = None, bias: torch.Tensor | None = None, src2: torch.Tensor | None = None) -> None:
    """dst[..., off:off+c_valid] = relu(src[..., :c_valid] (+ src2) + bias) for dense channels-last buffers [..., C]."""
    C = src.shape[-1]
    npix = src.numel() // C
    check(load().sdof_relu_scatter(ptr(src), ptr(src2), ptr(bias), npix, C, ptr(dst1), dst1.shape[-1], off1, ptr(dst2),
                                   dst2.shape[-1] if dst2 is not None else 0, off2, C if c_valid is None else c_valid,
                                   stream_ptr(src.device)), 'sdof_relu_scatter')


def _bias_is_map(bias, width: int) -> int:
    """per-channel vector [width] -> 0, per-pixel map [..., width] -> 1."""
    return int(bias is not None and bias.dim() > 1)


def gru_rh(zr: torch.Tensor, h: torch.Tensor, rhx: torch.Tensor, bias_zr: torch.Tensor | None = None) -> None:
    hidden = h.shape[-1]
    check(load().sdof_gru_rh(ptr(zr), ptr(bias_zr), ptr(h), ptr(rhx), h.numel() // hidden, hidden, rhx.shape[-1],
                             _bias_is_map(bias_zr, 2 * hidden), zr.shape[-1], stream_ptr(h.device)), 'sdof_gru_rh')


def gru_update(zr: torch.Tensor, q: torch.Tensor, h: torch.Tensor, hx: torch.Tensor, bias_zr: torch.Tensor | None = None,
               bias_q: torch.Tensor | None = None) -> None:
    hidden = h.shape[-1]
    if _bias_is_map(bias_zr, 2 * hidden) != _bias_is_map(bias_q, hidden):
        raise RuntimeError('gru_update: bias_zr and bias_q must both be vectors or both be per-pixel maps')
    check(load().sdof_gru_update(ptr(zr), ptr(bias_zr), ptr(q), ptr(bias_q), ptr(h), ptr(hx), h.numel() // hidden, hidden,
                                 hx.shape[-1], _bias_is_map(bias_zr, 2 * hidden), zr.shape[-1], stream_ptr(h.device)),
          'sdof_gru_update')


f16 = torch.float16


def gru_weights16(w: torch.Tensor) -> torch.Tensor:
    """Conv weight [Cout, Cin, 1, 5] or [Cout, Cin, 5, 1] -> K-major fp16 [Cout, 5*Cin] (k = tap*Cin + c) for the tcgen05 GRU."""
    co, ci, kh, kw = w.shape
    if (kh, kw) not in ((1, 5), (5, 1)):
        raise RuntimeError(f'expected a 1x5 or 5x1 kernel, got {kh}x{kw}')
    taps = w.reshape(co, ci, 5).permute(0, 2, 1)
    return taps.reshape(co, 5 * ci).to(f16).contiguous()


def motion_tail16(mc: torch.Tensor, mf: torch.Tensor, bias: torch.Tensor, flow: torch.Tensor, hx16: torch.Tensor) -> None:
    """hx16[..., 128:254] = relu(mc + mf + bias)[..., :126]; hx16[..., 254:256] = flow (fp16 GRU input, csrc/conv_tc.cu)."""
    npix = flow.numel() // 2
    check(load().sdof_motion_tail16(ptr(mc), ptr(mf), ptr(bias), ptr(flow), npix, ptr(hx16), hx16.shape[-1], stream_ptr(hx16.device)),
          'sdof_motion_tail16')


def gru_zr_tc(hx16: torch.Tensor, w_zr16: torch.Tensor, zrmap: torch.Tensor, h: torch.Tensor, horizontal: bool, z: torch.Tensor,
              rh16: torch.Tensor, qx: torch.Tensor) -> None:
    """z | r | q_x convolution of one SepConvGRU pass on tcgen05 with the gate arithmetic fused (sdof_gru_zr_tc)."""
    B, hh, ww, C = hx16.shape
    if C != 256 or hx16.dtype != f16 or tuple(w_zr16.shape) != (384, 5 * 256) or w_zr16.dtype != f16:
        raise RuntimeError('gru_zr_tc: hx16 must be fp16 [B,h,w,256] and w_zr16 fp16 [384, 1280]')
    check(load().sdof_gru_zr_tc(ptr(hx16), ptr(w_zr16), ptr(zrmap), ptr(h), B, hh, ww, int(horizontal), ptr(z), ptr(rh16), ptr(qx),
                                stream_ptr(hx16.device)), 'sdof_gru_zr_tc')


def gru_q_tc(rh16: torch.Tensor, w_q16: torch.Tensor, qmap: torch.Tensor, qx: torch.Tensor, z: torch.Tensor, horizontal: bool,
             h: torch.Tensor, hx16: torch.Tensor) -> None:
    """q convolution + hidden-state update of one SepConvGRU pass on tcgen05 (sdof_gru_q_tc); h is updated in place."""
    B, hh, ww, C = rh16.shape
    if C != 128 or rh16.dtype != f16 or tuple(w_q16.shape) != (128, 5 * 128) or w_q16.dtype != f16:
        raise RuntimeError('gru_q_tc: rh16 must be fp16 [B,h,w,128] and w_q16 fp16 [128, 640]')
    check(load().sdof_gru_q_tc(ptr(rh16), ptr(w_q16), ptr(qmap), ptr(qx), ptr(z), B, hh, ww, int(horizontal), ptr(h), ptr(hx16),
                               hx16.shape[-1], stream_ptr(rh16.device)), 'sdof_gru_q_tc')


# ---- fp16-activation forms (csrc/raft_glue16.cu) for the update block run with cuDNN fp16 convolutions
def corr_lookup_nhwc_h(pyr: 'CorrPyramid', coords_nhwc: torch.Tensor, out16: torch.Tensor) -> torch.Tensor:
    """coords [B,h1,w1,2] -> out16 [B,h1,w1,Cpad] fp16, Cpad >= 324 a multiple of 8 (padding channels zeroed); radius 4, 4 levels."""
    check(load().sdof_corr_lookup_h(ptr(pyr.buf), pyr.elem_bytes, ptr(coords_nhwc), pyr.B, pyr.h1, pyr.w1, pyr.h2, pyr.w2, pyr.levels, 4,
                                    ptr(out16), out16.shape[-1], stream_ptr(out16.device)), 'sdof_corr_lookup_h')
    return out16


def corr_lookup_gather_nhwc_h(pyr: 'CorrPyramid', coords_in: torch.Tensor, taps: torch.Tensor | None, tap_bias, coords_out: torch.Tensor,
                              flow_out: torch.Tensor, out16: torch.Tensor) -> torch.Tensor:
    """corr_lookup_nhwc_h on coords = coords_in + gather(taps) (the deferred coords update, see include/sdof_b200.h); writes the new
    coordinates to coords_out (a different tensor) and flow_out = coords_out - grid.  taps None: coordinates pass through."""
    if coords_out.data_ptr() == coords_in.data_ptr():
        raise RuntimeError('corr_lookup_gather_nhwc_h: coords_out must not alias coords_in')
    check(load().sdof_corr_lookup_gather_h(ptr(pyr.buf), pyr.elem_bytes, ptr(coords_in), ptr(taps), float(tap_bias[0]), float(tap_bias[1]),
                                           ptr(coords_out), ptr(flow_out), pyr.B, pyr.h1, pyr.w1, pyr.h2, pyr.w2, pyr.levels, 4, ptr(out16),
                                           out16.shape[-1], stream_ptr(out16.device)), 'sdof_corr_lookup_gather_h')
    return out16


def conv7x7_c2_relu_coords_h(coords: torch.Tensor, taps: torch.Tensor | None, tap_bias, wT: torch.Tensor, bias: torch.Tensor,
                             out16: torch.Tensor) -> torch.Tensor:
    """conv7x7_c2_relu_h on flow = coords + gather(taps) - grid, computed on the fly from coords [B,h,w,2] and the taps."""
    B, h, w, _ = coords.shape
    check(load().sdof_conv7x7_c2_relu_coords_h(ptr(coords), ptr(taps), float(tap_bias[0]), float(tap_bias[1]), ptr(wT), ptr(bias), ptr(out16),
                                               B, h, w, stream_ptr(out16.device)), 'sdof_conv7x7_c2_relu_coords_h')
    return out16


def flow_im2col7_h(coords: torch.Tensor, taps: torch.Tensor | None, tap_bias, out16: torch.Tensor) -> torch.Tensor:
    """im2col rows of the 7x7 x 2-channel convolution of flow = coords + gather(taps) - grid, fp16 hi/lo split (include/sdof_b200.h):
    out16 [B,h,w,Kpad] with Kpad >= 196 a multiple of 8.  `im2col7_weight` builds the matching 1x1 filter."""
    B, h, w, _ = coords.shape
    check(load().sdof_flow_im2col7_h(ptr(coords), ptr(taps), float(tap_bias[0]), float(tap_bias[1]), ptr(out16), out16.shape[-1], B, h, w,
                                     stream_ptr(out16.device)), 'sdof_flow_im2col7_h')
    return out16


def im2col7_weight(weight: torch.Tensor, kpad: int = 200) -> torch.Tensor:
    """convf1's filter [Cout,2,7,7] -> the fp16 1x1 filter [Cout,kpad,1,1] (channels-last) over flow_im2col7_h's rows: [tap][ci] for
    the hi half, the same again for the lo half, zeros for the padding."""
    co = weight.shape[0]
    wk = weight.detach().float().permute(0, 2, 3, 1).reshape(co, 98)            # [Cout][(ky*7+kx)*2 + ci]
    full = torch.zeros((co, kpad), device=weight.device, dtype=torch.float32)
    full[:, :98] = wk
    full[:, 98:196] = wk
    return full.half().view(co, kpad, 1, 1).contiguous(memory_format=torch.channels_last)


def conv7x7_c2_relu_h(flow_nhwc: torch.Tensor, wT: torch.Tensor, bias: torch.Tensor, out16: torch.Tensor) -> torch.Tensor:
    B, h, w, _ = flow_nhwc.shape
    check(load().sdof_conv7x7_c2_relu_h(ptr(flow_nhwc), ptr(wT), ptr(bias), ptr(out16), B, h, w, stream_ptr(out16.device)),
          'sdof_conv7x7_c2_relu_h')
    return out16


def motion_tail16_h(mc16: torch.Tensor, mf16: torch.Tensor, bias: torch.Tensor, flow: torch.Tensor, hx16: torch.Tensor) -> None:
    check(load().sdof_motion_tail16_h(ptr(mc16), ptr(mf16), ptr(bias), ptr(flow), flow.numel() // 2, ptr(hx16), hx16.shape[-1],
                                      stream_ptr(hx16.device)), 'sdof_motion_tail16_h')


def gru_rh_h(zr16: torch.Tensor, zrmap: torch.Tensor, h: torch.Tensor, rh16: torch.Tensor) -> None:
    check(load().sdof_gru_rh_h(ptr(zr16), zr16.shape[-1], ptr(zrmap), ptr(h), ptr(rh16), h.numel() // 128, stream_ptr(h.device)),
          'sdof_gru_rh_h')


def gru_update_h(zr16: torch.Tensor, zrmap: torch.Tensor, q16: torch.Tensor, qmap: torch.Tensor, h: torch.Tensor, hx16: torch.Tensor,
                 h16: torch.Tensor | None) -> None:
    if zr16.shape[-1] != 384:
        raise RuntimeError('gru_update_h: zr16 must carry [z | r | q_x] = 384 channels')
    check(load().sdof_gru_update_h(ptr(zr16), ptr(zrmap), ptr(q16), ptr(qmap), ptr(h), ptr(hx16), hx16.shape[-1], ptr(h16),
                                   h.numel() // 128, stream_ptr(h.device)), 'sdof_gru_update_h')


def flowhead2_update_h(x16: torch.Tensor, w2: torch.Tensor, bias, coords1: torch.Tensor, flow: torch.Tensor, scratch: torch.Tensor) -> None:
    """flowhead2_update on fp16 activations x16 [B,h,w,256]: tap products, then the 9-neighbour gather + coords / flow update."""
    B, h, w, C = x16.shape
    if C != 256 or x16.dtype != f16:
        raise RuntimeError('flowhead2_update_h: x16 must be fp16 [B,h,w,256]')
    lib = load()
    check(lib.sdof_flowhead2_taps_h(ptr(x16), ptr(w2), B * h * w, ptr(scratch), stream_ptr(x16.device)), 'sdof_flowhead2_taps_h')
    check(lib.sdof_flowhead2_gather_update(ptr(scratch), float(bias[0]), float(bias[1]), ptr(coords1), ptr(flow), None, 0, 0, B, h, w,
                                           stream_ptr(x16.device)), 'sdof_flowhead2_gather_update')


def flowhead2_taps_h(x16: torch.Tensor, w2: torch.Tensor, scratch: torch.Tensor) -> None:
    """First half of flowhead2_update_h alone: the 18 tap products per pixel into scratch [B*h*w*18] (the consumers of the
    coordinates apply them: corr_lookup_gather_nhwc_h, conv7x7_c2_relu_coords_h, and flowhead2_gather_update after the loop)."""
    B, h, w, C = x16.shape
    if C != 256 or x16.dtype != f16:
        raise RuntimeError('flowhead2_taps_h: x16 must be fp16 [B,h,w,256]')
    check(load().sdof_flowhead2_taps_h(ptr(x16), ptr(w2), B * h * w, ptr(scratch), stream_ptr(x16.device)), 'sdof_flowhead2_taps_h')


def flowhead2_gather_update(scratch: torch.Tensor, bias, coords1: torch.Tensor, flow: torch.Tensor) -> None:
    """coords1 += bias + 9-neighbour sum of the taps in scratch (in place); flow = coords1 - grid."""
    B, h, w, _ = coords1.shape
    check(load().sdof_flowhead2_gather_update(ptr(scratch), float(bias[0]), float(bias[1]), ptr(coords1), ptr(flow), None, 0, 0, B, h, w,
                                              stream_ptr(coords1.device)), 'sdof_flowhead2_gather_update')


def flow_update(delta: torch.Tensor | None, coords1: torch.Tensor, flow: torch.Tensor, hx: torch.Tensor | None, hx_off: int,
                rhx: torch.Tensor | None, rhx_off: int, delta_bias=(0.0, 0.0)) -> None:
    B, h, w, _ = coords1.shape
    check(load().sdof_flow_update(ptr(delta), float(delta_bias[0]), float(delta_bias[1]), ptr(coords1), ptr(flow), ptr(hx),
                                  hx.shape[-1] if hx is not None else 0, hx_off, ptr(rhx),
                                  rhx.shape[-1] if rhx is not None else 0, rhx_off, B, h, w, stream_ptr(coords1.device)),
          'sdof_flow_update')


def convex_upsample(mask_nhwc: torch.Tensor, flow_nhwc: torch.Tensor, mask_scale: float = 0.25,
                    mask_bias: torch.Tensor | None = None) -> torch.Tensor:
    """mask [B,h,w,576] (+ bias[576]), flow [B,h,w,2] -> [B,8h,8w,2] (RAFT.upsample_flow, raft.py:72-83)."""
    require_cuda(mask_nhwc, 'mask', f32)
    require_cuda(flow_nhwc, 'flow', f32)
    B, h, w, _ = flow_nhwc.shape
    if tuple(mask_nhwc.shape) != (B, h, w, 576):
        raise RuntimeError(f'mask must be {(B, h, w, 576)}, got {tuple(mask_nhwc.shape)}')
    up = torch.empty((B, 8 * h, 8 * w, 2), dtype=f32, device=flow_nhwc.device)
    check(load().sdof_convex_upsample(ptr(mask_nhwc), ptr(mask_bias), float(mask_scale), ptr(flow_nhwc), B, h, w, ptr(up),
                                      stream_ptr(up.device)), 'sdof_convex_upsample')
    return up


def instnorm_relu(x: torch.Tensor, relu: bool = True, eps: float = 1e-5, inplace: bool = True) -> torch.Tensor:
    """InstanceNorm2d(no affine) + optional ReLU on a contiguous NCHW tensor."""
    require_cuda(x, 'x', f32)
    N, C, H, W = x.shape
    y = x if inplace else torch.empty_like(x)
    check(load().sdof_instnorm_relu_nchw(ptr(x), ptr(y), N * C, H * W, float(eps), int(relu), stream_ptr(x.device)),
          'sdof_instnorm_relu_nchw')
    return y


def _nhwc_dims(x: torch.Tensor):
    """x: 4-D tensor [N,C,H,W] stored channels-last (dense NHWC)."""
    require_cuda_cl(x, 'x')
    N, C, H, W = x.shape
    return N, C, H * W


def require_cuda_cl(t: torch.Tensor, name: str, dtypes=(f32,)) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype not in dtypes or t.dim() != 4:
        raise RuntimeError(f'{name} must be a 4-D CUDA tensor of dtype {" / ".join(str(d) for d in dtypes)}')
    if not t.is_contiguous(memory_format=torch.channels_last):
        raise RuntimeError(f'{name} must be dense channels_last (NHWC)')


def instnorm_nhwc(x: torch.Tensor, stats: torch.Tensor, relu: bool = True, residual: torch.Tensor | None = None,
                  eps: float = 1e-5) -> torch.Tensor:
    """In-place InstanceNorm2d(no affine) (+ReLU) (+ `relu(residual + .)`) on a channels-last [N,C,H,W] tensor, fp32 or fp16
    (statistics always accumulate in fp32 partials / fp64 atomics).
    stats: zeroed fp64 scratch with at least N*C*2 elements (consumed by this call)."""
    require_cuda_cl(x, 'x', (f32, torch.float16))
    N, C, H, W = x.shape
    hw = H * W
    if residual is not None:
        require_cuda_cl(residual, 'residual', (x.dtype,))
        if residual.shape != x.shape:
            raise RuntimeError('residual must have the shape of x')
    if stats.dtype != torch.float64 or not stats.is_cuda or stats.numel() < N * C * 2 or not stats.is_contiguous():
        raise RuntimeError('stats must be a contiguous fp64 CUDA tensor with >= N*C*2 elements')
    lib = load()
    if x.dtype == f32:
        check(lib.sdof_instnorm_stats_nhwc(ptr(x), N, hw, C, ptr(stats), stream_ptr(x.device)), 'sdof_instnorm_stats_nhwc')
        check(lib.sdof_instnorm_apply_nhwc(ptr(x), ptr(stats), ptr(residual), ptr(x), N, hw, C, float(eps), int(relu),
                                           stream_ptr(x.device)), 'sdof_instnorm_apply_nhwc')
    else:
        check(lib.sdof_instnorm_stats_nhwc_h(ptr(x), N, hw, C, ptr(stats), stream_ptr(x.device)), 'sdof_instnorm_stats_nhwc_h')
        check(lib.sdof_instnorm_apply_nhwc_h(ptr(x), ptr(stats), ptr(residual), ptr(x), N, hw, C, float(eps), int(relu),
                                             stream_ptr(x.device)), 'sdof_instnorm_apply_nhwc_h')
    return x


def add_relu_(y: torch.Tensor, a: torch.Tensor) -> torch.Tensor:
    """y = relu(a + y) in place; both dense with identical layout."""
    if y.shape != a.shape or y.stride() != a.stride() or y.dtype != f32 or a.dtype != f32 or not y.is_cuda or not a.is_cuda:
        raise RuntimeError('add_relu_: tensors must be fp32 CUDA tensors of identical shape and layout')
    check(load().sdof_add_relu(ptr(a), ptr(y), ptr(y), y.numel(), stream_ptr(y.device)), 'sdof_add_relu')
    return y


def conv7x7_c2_relu(flow_nhwc: torch.Tensor, wT: torch.Tensor, bias: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """relu(conv7x7(flow) + bias): flow [B,h,w,2] -> [B,h,w,128]; wT = convf1.weight.permute(2,3,1,0) contiguous."""
    require_cuda(flow_nhwc, 'flow', f32)
    B, h, w, _ = flow_nhwc.shape
    if tuple(wT.shape) != (7, 7, 2, 128) or not wT.is_contiguous():
        raise RuntimeError(f'wT must be contiguous [7,7,2,128], got {tuple(wT.shape)}')
    if out is None:
        out = torch.empty((B, h, w, 128), dtype=f32, device=flow_nhwc.device)
    check(load().sdof_conv7x7_c2_relu(ptr(flow_nhwc), ptr(wT), ptr(bias), ptr(out), B, h, w, stream_ptr(out.device)),
          'sdof_conv7x7_c2_relu')
    return out


def flowhead2_update(x_nhwc: torch.Tensor, w2: torch.Tensor, bias, coords1: torch.Tensor, flow: torch.Tensor,
                     hx: torch.Tensor | None, hx_off: int, rhx: torch.Tensor | None, rhx_off: int,
                     scratch: torch.Tensor | None = None) -> None:
    """delta = conv3x3(x) + bias; coords1 += delta; flow = coords1 - grid (into flow and the hx / rhx flow slots)."""
    require_cuda(x_nhwc, 'x', f32)
    B, h, w, C = x_nhwc.shape
    if C != 256 or tuple(w2.shape) != (3, 3, 2, 256) or not w2.is_contiguous():
        raise RuntimeError('flowhead2_update: x must be [B,h,w,256] and w2 contiguous [3,3,2,256]')
    if scratch is None:
        scratch = torch.empty((B * h * w * 18,), dtype=f32, device=x_nhwc.device)
    elif scratch.numel() < B * h * w * 18 or scratch.dtype != f32 or not scratch.is_cuda:
        raise RuntimeError('flowhead2_update: scratch must be an fp32 CUDA tensor with >= B*h*w*18 elements')
    check(load().sdof_flowhead2_update(ptr(x_nhwc), ptr(w2), float(bias[0]), float(bias[1]), ptr(coords1), ptr(flow), ptr(hx),
                                       hx.shape[-1] if hx is not None else 0, hx_off, ptr(rhx),
                                       rhx.shape[-1] if rhx is not None else 0, rhx_off, B, h, w, ptr(scratch),
                                       stream_ptr(x_nhwc.device)),
          'sdof_flowhead2_update')


# --------------------------------------------------------------------------- after the path (guided_ldm_inpainting.py:290-309)
def mask_blur_composite(mask: torch.Tensor, image: torch.Tensor | None, reference: torch.Tensor | None, mask_blur: float):
    """blurred = GaussianBlur(mask_blur)(mask); out = Image.composite(reference, image, blurred), bit-exact to Pillow.
    mask u8 [B,H,W]; image, reference u8 [B,H,W,C] or both None (blur only).  Returns (out or None, blurred)."""
    require_cuda(mask, 'mask', u8)
    B, H, W = mask.shape
    out, C = None, 1
    if image is not None or reference is not None:
        require_cuda(image, 'image', u8)
        require_cuda(reference, 'reference', u8)
        if image.shape != reference.shape or tuple(image.shape[:3]) != (B, H, W):
            raise RuntimeError('image and reference must both be [B,H,W,C] matching the mask')
        C = image.shape[3]
        out = torch.empty_like(image)
    blurred = torch.empty_like(mask)
    if B == 0:
        return out, blurred
    check(load().sdof_mask_blur_composite(ptr(mask), ptr(image), ptr(reference), B, H, W, C, float(mask_blur), ptr(blurred), ptr(out),
                                          stream_ptr(mask.device)), 'sdof_mask_blur_composite')
    return out, blurred


def resize_bicubic_u8(src: torch.Tensor, oh: int, ow: int, want_latmask: bool = False):
    """Image.resize((ow, oh)) (BICUBIC) of u8 [B,H,W], bit-exact to Pillow.  Returns dst u8 [B,oh,ow] and, with
    want_latmask, also f32 [B,4,oh,ow] = around(dst / 255) tiled over 4 channels (guided_ldm_inpainting.py:304-308)."""
    require_cuda(src, 'src', u8)
    B, H, W = src.shape
    nbytes = int(load().sdof_resize_bicubic_workspace_bytes(B, H, W, oh, ow))
    if nbytes < 0:
        raise RuntimeError('resize_bicubic_u8: bad sizes')
    ws = torch.empty((nbytes,), dtype=u8, device=src.device)
    dst = torch.empty((B, oh, ow), dtype=u8, device=src.device)
    lat = torch.empty((B, 4, oh, ow), dtype=f32, device=src.device) if want_latmask else None
    check(load().sdof_resize_bicubic_u8(ptr(src), B, H, W, oh, ow, ptr(dst), ptr(lat), ptr(ws), nbytes, stream_ptr(src.device)),
          'sdof_resize_bicubic_u8')
    return (dst, lat) if want_latmask else dst


# --------------------------------------------------------------------------- before the path (ofgen_pixel_inpaint.py:127-176)
def detect_edges(frame_bgr: torch.Tensor, dilate_k: int, low: int = -1, high: int = -1) -> torch.Tensor:
    """cv2.dilate(cv2.Canny(V(frame), low, high), ones(k,k)); low/high < 0 = the reference's median rule.
    frame_bgr u8 [H,W,3] -> u8 [H,W] in {0,255}, bit-exact to OpenCV."""
    require_cuda(frame_bgr, 'frame', u8)
    if frame_bgr.dim() != 3 or frame_bgr.shape[2] != 3:
        raise RuntimeError(f'frame must be [H,W,3], got {tuple(frame_bgr.shape)}')
    H, W, _ = frame_bgr.shape
    nbytes = int(load().sdof_detect_edges_workspace_bytes(H, W))
    ws = torch.empty((nbytes,), dtype=u8, device=frame_bgr.device)
    edges = torch.empty((H, W), dtype=u8, device=frame_bgr.device)
    check(load().sdof_detect_edges(ptr(frame_bgr), H, W, int(dilate_k), int(low), int(high), ptr(edges), ptr(ws), nbytes,
                                   stream_ptr(frame_bgr.device)), 'sdof_detect_edges')
    return edges


def abs_diff_sum(a: torch.Tensor, b: torch.Tensor) -> int:
    """sum |a - b| over two u8 tensors of the same shape (exact integer)."""
    require_cuda(a, 'a', u8)
    require_cuda(b, 'b', u8)
    if a.shape != b.shape:
        raise RuntimeError('abs_diff_sum: shapes differ')
    out = torch.empty((1,), dtype=torch.int64, device=a.device)
    check(load().sdof_abs_diff_sum_u8(ptr(a), ptr(b), a.numel(), ptr(out), stream_ptr(a.device)), 'sdof_abs_diff_sum_u8')
    return int(out.item())


def normalize_pad_u8(img: torch.Tensor, pad, channels: int = 3, bgr: bool = False) -> torch.Tensor:
    """u8 [B,H,W,3] -> channels-last fp32 [B,channels,Hp,Wp] = 2*(x/255)-1 of the replicate-padded frame (channels = 4
    appends a zero channel; bgr=True reads BGR frames as RGB).  pad = (left, right, top, bottom) as in F.pad / InputPadder."""
    require_cuda(img, 'img', u8)
    B, H, W, C = img.shape
    if C != 3:
        raise RuntimeError('normalize_pad_u8: 3-channel frames expected')
    left, right, top, bottom = (int(v) for v in pad)
    Hp, Wp = H + top + bottom, W + left + right
    out = torch.empty((B, Hp, Wp, channels), dtype=f32, device=img.device)
    check(load().sdof_normalize_pad_u8_nhwc(ptr(img), B, H, W, top, left, Hp, Wp, channels, int(bgr), ptr(out), stream_ptr(img.device)),
          'sdof_normalize_pad_u8_nhwc')
    return out.permute(0, 3, 1, 2)
