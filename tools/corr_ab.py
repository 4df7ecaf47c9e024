"""A/B timing of the correlation-pyramid kernel between two builds of libsdof_b200.so ON THE SAME BOX (boxes differ by several
per cent): python tools/corr_ab.py <other_lib.so>... [rounds].  Alternates the two libraries in fresh processes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, torch
sys.path.insert(0, %r)
from sd_animation_optical_flow_b200 import ops
dev = torch.device('cuda', 0)
out = {}
for (h, w) in ((96, 64), (90, 160)):
    g = torch.Generator(device=dev).manual_seed(0)
    f1 = torch.randn((1, h, w, 256), generator=g, device=dev)
    f2 = torch.randn((1, h, w, 256), generator=g, device=dev)
    src, tgt = ops.CorrSource(f1, 'fp16'), ops.CorrTarget(f2, 4, 'fp16')
    pyr = src.pyramid(tgt, 'fp16')
    for _ in range(10):
        src.pyramid(tgt, 'fp16', out=pyr)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            src.pyramid(tgt, 'fp16', out=pyr)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 20 * 1e3)
    out[f'{h}x{w}'] = round(best, 2)
print(json.dumps(out))
''' % ROOT

others = [a for a in sys.argv[1:] if not a.isdigit()]
rounds = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 3
for r in range(rounds):
    for name, lib in [(os.path.basename(o), os.path.abspath(o)) for o in others] + [('tree', '')]:
        env = dict(os.environ, SDOF_B200_LIB=lib)
        res = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
        print(f'{name:28s}', res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-300:], flush=True)
