"""Fixed vs per-iteration cost of the headline step in its CUDA graph: time(iters) for several iteration counts and a linear fit.
    python tools/step_split.py"""
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
sty = torch.randint(0, 256, (1, 768, 512, 3), dtype=torch.uint8, device=dev, generator=g)
ks = (1, 4, 8, 12, 16, 20)
engines = {k: RaftEngine(iters=k, device=dev, flow_head_scale=0.02) for k in ks}
for e in engines.values():
    for _ in range(5):
        ops.warp(sty, e.estimate_flow(a, b), 'cv2_cubic', -1.0)
torch.cuda.synchronize()
t = {k: [] for k in ks}
for r in range(7):
    for k, e in engines.items():
        s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(100):
            ops.warp(sty, e.estimate_flow(a, b), 'cv2_cubic', -1.0)
        f.record()
        torch.cuda.synchronize()
        t[k].append(s.elapsed_time(f) / 100)
med = {k: statistics.median(v) for k, v in t.items()}
n = len(ks)
mx, my = sum(ks) / n, sum(med.values()) / n
slope = sum((k - mx) * (med[k] - my) for k in ks) / sum((k - mx) ** 2 for k in ks)
print(json.dumps({'ms_by_iters': {k: round(v, 4) for k, v in med.items()}, 'per_iteration_us': round(slope * 1e3, 2),
                  'fixed_us (encoders, volume, upsample, warp, copies)': round((my - slope * mx) * 1e3, 1)}))
