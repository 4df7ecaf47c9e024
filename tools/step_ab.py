"""In-process A/B of two FastRaft option sets on the headline step (one 768x512 pair, iters 20, CUDA graph): the variants are timed
alternately in the same process so that box-to-box and run-to-run drift (~2-3 % between `bench.py --quick` invocations) cancels.
    python tools/step_ab.py '{"defer_coords": true}' '{"defer_coords": false}' [rounds]"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sd_animation_optical_flow_b200 import ops  # noqa: E402
from sd_animation_optical_flow_b200.engine import RaftEngine  # noqa: E402

opts = [json.loads(a) for a in sys.argv[1:3]]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 7
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randint(0, 256, (1, 768, 512, 3), dtype=torch.uint8, device=dev, generator=g)
b = a.roll(3, 1)
sty = torch.randint(0, 256, (1, 768, 512, 3), dtype=torch.uint8, device=dev, generator=g)
# an option set may carry "_env": {"SDOF_...": "..."}: set while THAT engine is built, warmed up and captured (launch-time switches
# of the library are baked into its CUDA graph)
engines = []
for o in opts:
    env = o.pop('_env', {}) if isinstance(o, dict) else {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    e = RaftEngine(iters=20, device=dev, flow_head_scale=0.02, fast_options=o)
    for _ in range(5):
        ops.warp(sty, e.estimate_flow(a, b), 'cv2_cubic', -1.0)
    torch.cuda.synchronize()
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    engines.append(e)
    o['_env'] = env
torch.cuda.synchronize()
times = [[], []]
for r in range(rounds):
    for i, e in enumerate(engines):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(150):
            ops.warp(sty, e.estimate_flow(a, b), 'cv2_cubic', -1.0)
        t.record()
        torch.cuda.synchronize()
        times[i].append(s.elapsed_time(t) / 150)
for o, t in zip(opts, times):
    print(json.dumps({'options': o, 'ms_per_step_median': round(statistics.median(t), 4), 'min': round(min(t), 4), 'max': round(max(t), 4)}))
