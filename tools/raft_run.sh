#!/bin/bash
mkdir -p gpurun_out
echo "=== raft perf"; timeout 600 python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
from sd_animation_optical_flow_b200.engine import RaftEngine
dev = torch.device('cuda', 0)
img = torch.randint(0, 256, (1, 768, 512, 3), dtype=torch.uint8, device=dev)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for bench_mode in (False, True):
  torch.backends.cudnn.benchmark = bench_mode
  for kw in (dict(fast=True, use_cuda_graph=True),):
    eng = RaftEngine(iters=20, device=dev, **kw)
    eng.estimate_flow(img, img.flip(1)); torch.cuda.synchronize()
    print('cudnn.benchmark', bench_mode, kw, f'{timeit(lambda: eng.estimate_flow(img, img.flip(1))):.2f} ms/pair', 'fused conv+relu:', getattr(eng.fast, '_fused_relu_ok', None) if eng.fast else None, flush=True)
PY
