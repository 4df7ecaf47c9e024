"""Breaks the e2e step of bench.py (numpy API, pinned host buffers) into its two calls.  python tools/e2e_probe.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sd_animation_optical_flow_b200 import ofgen  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402

H, W = 768, 512
f1, f2 = gi.shifted_pair(H, W, 1000)
sty = gi.texture(H, W, 2000)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
b1, b2, bs = pin(f1[:, :, ::-1]), pin(f2[:, :, ::-1]), pin(sty[:, :, ::-1])
algo = ofgen.RAFT_2(model_path=None, iters=20, use_cuda_graph=True)
for _ in range(3):
    ofgen.warp_frame(bs, algo.calc(b1, b2))
for rep in range(3):
    tc = tw = 0.0
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        ta = time.perf_counter()
        flow = algo.calc(b1, b2)
        tb = time.perf_counter()
        out = ofgen.warp_frame(bs, flow)
        tc += tb - ta
        tw += time.perf_counter() - tb
    dt = time.perf_counter() - t0
    print(f'rep {rep}: {n / dt:6.1f} pairs/s  calc {tc / n * 1e3:.3f} ms  warp_frame {tw / n * 1e3:.3f} ms')
# device-only reference for the same engine
a = torch.from_numpy(f1).cuda()[None]
b = torch.from_numpy(f2).cuda()[None]
eng = algo.engine
for _ in range(3):
    eng.estimate_flow(a, b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
    eng.estimate_flow(a, b)
torch.cuda.synchronize()
print(f'device-resident estimate_flow: {(time.perf_counter() - t0) / 30 * 1e3:.3f} ms')
