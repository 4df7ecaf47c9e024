"""`corr_fn` objects with the reference's protocol (RAFT/core/raft.py:104-107, 124):

    corr_fn = CorrBlock(fmap1[B,C,h,w], fmap2[B,C,h,w], num_levels=4, radius=4)
    corr    = corr_fn(coords1[B,2,h,w])  ->  fp32 [B, num_levels*(2r+1)^2, h, w], contiguous

`CorrBlock` replaces RAFT/core/corr.py:12-60 (matmul + scale + 3 avg_pool2d at build time,
4 grid_sample + cat + permute per call) with ONE tcgen05/TMA kernel at build time and ONE
lookup kernel per call.  `AlternateCorrBlock` replaces RAFT/core/corr.py:63-91 and never
materialises the volume.  Both run only on CUDA tensors and raise otherwise.
"""
from __future__ import annotations

import math

import torch

from . import ops


def _to_nhwc(fmap: torch.Tensor) -> torch.Tensor:
    """[B,C,h,w] (any memory format) -> contiguous [B,h,w,C] fp32.  A channels_last
    feature map is already in that layout, so this is free for it."""
    if fmap.dim() != 4:
        raise RuntimeError(f'feature map must be [B,C,h,w], got {tuple(fmap.shape)}')
    return fmap.float().permute(0, 2, 3, 1).contiguous()


class CorrBlock:
    def __init__(self, fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, radius: int = 4,
                 precision: str = 'fp16'):
        if not fmap1.is_cuda or not fmap2.is_cuda:
            raise RuntimeError('CorrBlock needs CUDA feature maps: this package has no CPU path')
        self.num_levels = num_levels
        self.radius = radius
        self.precision = precision
        self.pyramid = ops.corr_volume_pyramid(_to_nhwc(fmap1), _to_nhwc(fmap2), num_levels, precision)

    @property
    def corr_pyramid(self):
        """List of [B*h*w, 1, h_l, w_l] views, the attribute the reference exposes (corr.py:15)."""
        return [self.pyramid.level(l) for l in range(self.num_levels)]

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        return ops.corr_lookup(self.pyramid, coords.float().contiguous(), self.radius)

    @staticmethod
    def corr(fmap1: torch.Tensor, fmap2: torch.Tensor, precision: str = 'fp16') -> torch.Tensor:
        """CorrBlock.corr (corr.py:52-60): [B,h,w,1,h,w] volume divided by sqrt(C)."""
        B, _, h, w = fmap1.shape
        pyr = ops.corr_volume_pyramid(_to_nhwc(fmap1), _to_nhwc(fmap2), 1, precision)
        return pyr.level(0).reshape(B, h, w, 1, fmap2.shape[2], fmap2.shape[3])


class AlternateCorrBlock:
    def __init__(self, fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, radius: int = 4):
        if not fmap1.is_cuda or not fmap2.is_cuda:
            raise RuntimeError('AlternateCorrBlock needs CUDA feature maps: this package has no CPU path')
        self.num_levels = num_levels
        self.radius = radius
        self.dim = fmap1.shape[1]
        self.fmap1 = _to_nhwc(fmap1)
        # only fmap2 is needed per level (the reference also pools fmap1 but never reads it, corr.py:82-83)
        self.fmap2_levels = [_to_nhwc(fmap2)]
        for _ in range(num_levels - 1):
            self.fmap2_levels.append(ops.avgpool2_nhwc(self.fmap2_levels[-1]))

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        coords = coords.float().contiguous()
        B, _, H, W = coords.shape
        dd = (2 * self.radius + 1) ** 2
        out = torch.empty((B, self.num_levels * dd, H, W), dtype=torch.float32, device=coords.device)
        inv = 1.0 / math.sqrt(self.dim)
        for i, f2 in enumerate(self.fmap2_levels):
            if f2.shape[1] == 0 or f2.shape[2] == 0:
                out[:, i * dd:(i + 1) * dd].zero_()
                continue
            ops.alt_corr_level(self.fmap1, f2, coords, self.radius, 1.0 / 2 ** i, inv, i * dd, out)
        return out
