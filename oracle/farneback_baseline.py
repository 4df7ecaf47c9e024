"""CPU baseline the north star names: cv2.calcOpticalFlowFarneback + torch.grid_sample on the host.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py): used by bench.py's `cpu_baseline` leg and
`--impl reference`, never by the product path.  Parameters are the ones left in the reference's
commented-out call (ofgen_keyframe_inpaint.py:134): (None, 0.5, 5, 15, 3, 5, 1.2, 0); the warp is the
bilinear backward warp of RAFT/core/utils/utils.py:57-71 (grid_sample, zeros, align_corners=True).
"""
from __future__ import annotations

import os
import time

import numpy as np


def farneback_flow(frame1_bgr: np.ndarray, frame2_bgr: np.ndarray) -> np.ndarray:
    import cv2
    g1 = cv2.cvtColor(frame1_bgr, cv2.COLOR_BGR2GRAY)
    g2 = cv2.cvtColor(frame2_bgr, cv2.COLOR_BGR2GRAY)
    return cv2.calcOpticalFlowFarneback(g1, g2, None, 0.5, 5, 15, 3, 5, 1.2, 0)


def grid_sample_warp(frame_bgr: np.ndarray, flow: np.ndarray) -> np.ndarray:
    import torch
    import torch.nn.functional as F
    H, W = flow.shape[:2]
    img = torch.from_numpy(frame_bgr).permute(2, 0, 1).float()[None]
    fl = torch.from_numpy(flow)
    xs = torch.arange(W).float()[None, :] + fl[..., 0]
    ys = torch.arange(H).float()[:, None] + fl[..., 1]
    grid = torch.stack([2 * xs / (W - 1) - 1, 2 * ys / (H - 1) - 1], -1)[None]
    out = F.grid_sample(img, grid, mode='bilinear', padding_mode='zeros', align_corners=True)
    return out[0].permute(1, 2, 0).numpy()


def flow_and_warp(frame1_bgr, frame2_bgr, stylised_bgr):
    flow = farneback_flow(frame1_bgr, frame2_bgr)
    return flow, grid_sample_warp(stylised_bgr, flow)


def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host thread it can."""
    import cv2
    import torch
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    torch.set_num_threads(n)
    cv2.setNumThreads(n)
    return n


def host_info():
    import cv2
    import torch
    model = ''
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                model = line.split(':', 1)[1].strip()
                break
    except OSError:
        pass
    return {'cores': os.cpu_count(), 'cpu_model': model, 'cv2_threads': cv2.getNumThreads(),
            'torch_threads': torch.get_num_threads()}


def time_pairs(frame1, frame2, stylised, budget_s: float = 10.0, min_pairs: int = 3, max_pairs: int = 1000):
    """Run flow+warp on the same pair until `budget_s` of CPU time is spent.  Returns
    (pairs_per_s, n_pairs, per-pair seconds list)."""
    flow_and_warp(frame1, frame2, stylised)  # warm-up (thread pools, page faults)
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < max_pairs and (len(times) < min_pairs or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        flow_and_warp(frame1, frame2, stylised)
        times.append(time.perf_counter() - t0)
    return len(times) / sum(times), len(times), times
