"""In-graph timings of the per-iteration glue kernels at batch-1 size.   python tools/glue_bench.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from sd_animation_optical_flow_b200 import ops  # noqa: E402
from gru_bench import time_graphed  # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(s, generator=g, device=dev)
B, h, w = 1, 96, 64
row = {}
x = torch.relu(rnd(B, h, w, 256))
w2 = (rnd(3, 3, 2, 256) * 0.05).contiguous()
c1 = rnd(B, h, w, 2).contiguous()
fl = torch.empty((B, h, w, 2), device=dev)
hx = torch.zeros((B, h, w, 256), device=dev)
scratch = torch.empty((B * h * w * 18,), device=dev)
row['flowhead2_update_us'] = time_graphed(lambda: ops.flowhead2_update(x, w2, (0.1, 0.2), c1, fl, hx, 254, None, 0, scratch=scratch))
flow = rnd(B, h, w, 2)
wt = (rnd(7, 7, 2, 128) * 0.1).contiguous()
b7 = rnd(128)
row['conv7x7_c2_relu_us'] = time_graphed(lambda: ops.conv7x7_c2_relu(flow, wt, b7))
from sd_animation_optical_flow_b200.raft import coords_grid
f1, f2 = rnd(1, h, w, 256), rnd(1, h, w, 256)
pyr = ops.corr_volume_pyramid(f1, f2, 4, 'fp16', 'fp16')
cn = (coords_grid(1, h, w, dev) + 2 * rnd(1, 2, h, w)).permute(0, 2, 3, 1).contiguous()
out = torch.empty((1, h, w, 324), device=dev)
row['lookup_nhwc_fp16_us'] = time_graphed(lambda: ops.corr_lookup_nhwc(pyr, cn, 4, out))
print(json.dumps(row))
