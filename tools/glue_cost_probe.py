"""Critical-path cost of each hand-written kernel of the update loop inside the CUDA graph: the headline step is re-captured with
ONE op replaced by a no-op (results are wrong by construction; timing only) and the difference to the full step is what the best
possible fusion / optimisation of that op could return.   python tools/glue_cost_probe.py"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sd_animation_optical_flow_b200 import ops  # noqa: E402
from sd_animation_optical_flow_b200.engine import RaftEngine  # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randint(0, 256, (1, 768, 512, 3), dtype=torch.uint8, device=dev, generator=g)
b = a.roll(3, 1)


def build():
    e = RaftEngine(iters=20, device=dev, flow_head_scale=0.02)
    for _ in range(4):
        e.estimate_flow(a, b)
    torch.cuda.synchronize()
    return e


names = ['corr_lookup_gather_nhwc_h', 'conv7x7_c2_relu_coords_h', 'motion_tail16_h', 'gru_rh_h', 'gru_update_h', 'flowhead2_taps_h', 'instnorm_nhwc',
         'convex_upsample']
engines = {'full': build()}
for n in names:
    orig = getattr(ops, n)
    if n == 'instnorm_nhwc':
        setattr(ops, n, lambda y, stats, relu=True, residual=None: y)
    elif n == 'convex_upsample':
        setattr(ops, n, lambda mask, flow, scale, mask_bias=None: torch.zeros((flow.shape[0], flow.shape[1] * 8, flow.shape[2] * 8, 2), device=flow.device))
    else:
        setattr(ops, n, lambda *a_, **k_: (a_[-1] if n in ('corr_lookup_gather_nhwc_h', 'conv7x7_c2_relu_coords_h') else None))
    engines[n] = build()
    setattr(ops, n, orig)
t = {k: [] for k in engines}
for r in range(7):
    for k, e in engines.items():
        s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(100):
            e.estimate_flow(a, b)
        f.record()
        torch.cuda.synchronize()
        t[k].append(s.elapsed_time(f) / 100)
med = {k: statistics.median(v) for k, v in t.items()}
print(json.dumps({'full_ms': round(med['full'], 4), 'saved_us_when_removed': {k: round((med['full'] - v) * 1e3, 1) for k, v in med.items() if k != 'full'}}))
