"""Drop-in for the reference's pybind module `alt_cuda_corr`
(RAFT/alt_cuda_corr/correlation.cpp:51-54; imported by RAFT/core/corr.py:5-9).

    import sd_animation_optical_flow_b200.alt_cuda_corr as alt_cuda_corr
    corr, = alt_cuda_corr.forward(fmap1, fmap2, coords, radius)

Same argument meaning, return shape ([B,N,(2r+1)^2,H1,W1], unnormalised, in a 1-element
list) and error behaviour (RuntimeError for non-CUDA / non-contiguous inputs,
correlation.cpp:19-21); unlike the reference it also rejects non-fp32 inputs up front, runs
on the CURRENT stream under a device guard, and checks the launch.
"""
from __future__ import annotations

import torch

from . import ops


def forward(fmap1: torch.Tensor, fmap2: torch.Tensor, coords: torch.Tensor, radius: int):
    for name, t in (('fmap1', fmap1), ('fmap2', fmap2), ('coords', coords)):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError(f'{name} must be a CUDA tensor')
    with torch.cuda.device(fmap1.device):
        return [ops.alt_corr_forward(fmap1, fmap2, coords, int(radius))]


def backward(fmap1, fmap2, coords, corr_grad, radius):
    """Training-only op of the reference (correlation_kernel.cu:122-324).  The ofgen scripts run
    inference under torch.no_grad (ofgen.py:70); out of scope here."""
    raise NotImplementedError('alt_cuda_corr.backward is training-only and out of scope for the inference hot path')
