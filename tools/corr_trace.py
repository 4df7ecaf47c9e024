"""Device-side timeline of corr_pyramid_resident_kernel (SDOF_RES_TRACE): where a tile's time goes, per warp role.
    SDOF_NVCC_EXTRA=-DSDOF_RES_TRACE python -m sd_animation_optical_flow_b200.build     # the stamps are compiled out of the product build
    python tools/corr_trace.py [S|L] [fp16|fp32] [debug bits]
Stamps are clock64 of the CTA's SM (see RES_TRACE in csrc/corr_tc_res.cu); printed in ns at the SM clock given by SM_MHZ."""
import os
import struct
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sd_animation_optical_flow_b200 import ops  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else 'S'
storage = sys.argv[2] if len(sys.argv) > 2 else 'fp16'
os.environ['SDOF_RES_DEBUG'] = sys.argv[3] if len(sys.argv) > 3 else '0'
mhz = float(os.environ.get('SM_MHZ', 1965))
h, w = (90, 160) if size == 'L' else (96, 64)
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn((1, h, w, 256), generator=g, device=dev)
f2 = torch.randn((1, h, w, 256), generator=g, device=dev)
src, tgt = ops.CorrSource(f1, 'fp16'), ops.CorrTarget(f2, 4, 'fp16')
pyr = src.pyramid(tgt, storage)
for _ in range(5):
    src.pyramid(tgt, storage, out=pyr)
torch.cuda.synchronize()
path = os.path.join(ROOT, 'gpurun_out', 'corr_trace.bin')
os.makedirs(os.path.dirname(path), exist_ok=True)
os.environ['SDOF_RES_TRACE'] = path
src.pyramid(tgt, storage, out=pyr)
torch.cuda.synchronize()
os.environ['SDOF_RES_TRACE'] = ''

raw = open(path, 'rb').read()
ctas, tiles, slots, total = struct.unpack('4q', raw[:32])
T = np.frombuffer(raw[32:], dtype=np.int64).reshape(-1, tiles, slots)[:ctas].astype(np.float64)
T[T == 0] = np.nan
# SM clock of this run: cycle counter against %globaltimer between the first and the last committed tile of every CTA
last = np.array([np.flatnonzero(np.isfinite(T[c, :, 2])).max() for c in range(ctas)])
dc = T[np.arange(ctas), last, 2] - T[:, 0, 2]
dg = T[np.arange(ctas), last, 11] - T[:, 0, 11]
if 'SM_MHZ' not in os.environ and np.nanmedian(dg) > 0:
    mhz = float(np.nanmedian(dc / dg) * 1e3)
ns = 1e3 / mhz
print(f'{size} {storage} debug={os.environ["SDOF_RES_DEBUG"]}: {ctas} CTAs, {total} tiles ({total / ctas:.1f} per CTA), {mhz:.0f} MHz (cycle counter vs globaltimer)')


def stat(name, x):
    x = x[np.isfinite(x)] * ns
    if x.size:
        print(f'  {name:58s} mean {x.mean():8.0f} ns   p10 {np.percentile(x, 10):8.0f}   p50 {np.percentile(x, 50):8.0f}   p90 {np.percentile(x, 90):8.0f}   n={x.size}')


C = T[:, :, :11]
t0 = np.nanmin(C, axis=(1, 2), keepdims=True)
end = np.nanmax(C, axis=(1, 2))
print(f'  CTA lifetime first stamp -> last stamp: mean {np.nanmean(end - t0[:, 0, 0]) * ns:.0f} ns, max {np.nanmax(end - t0[:, 0, 0]) * ns:.0f} ns')
s = lambda k: T[:, :, k]
mid = slice(2, None)  # steady state: skip the first two tiles of every CTA
stat('MMA thread: tile period (commit -> commit)', (s(2)[:, 1:] - s(2)[:, :-1])[:, 1:])
stat('MMA thread: blocked on the accumulator (tempty)', (s(0)[:, 1:] - s(2)[:, :-1])[:, 1:])
stat('MMA thread: wait for the first slab', (s(1) - s(0))[:, mid])
stat('MMA thread: slabs 1..3 + issue + commit', (s(2) - s(1))[:, mid])
stat('commit -> epilogue warp 0 sees the accumulator', (s(3) - s(2))[:, mid])
stat('epilogue: accumulator seen -> 64 columns in registers', (s(4) - s(3))[:, mid])
stat('epilogue: stores of columns 0-31 issued', (s(5) - s(4))[:, mid])
stat('epilogue: stores of columns 32-63 issued', (s(7) - s(5))[:, mid])
stat('epilogue: stores issued -> next accumulator seen', (s(3)[:, 1:] - s(7)[:, :-1])[:, 1:])
stat('epilogue: tile period', (s(4)[:, 1:] - s(4)[:, :-1])[:, 1:])
stat('producer: tile period (last slab issued)', (s(8)[:, 1:] - s(8)[:, :-1])[:, 1:])
stat('producer ahead of MMA commit (slab issue -> tile commit)', (s(2) - s(8))[:, mid])
stat('resident block: producer starts load -> MMA thread sees it', np.nanmax(s(10), axis=1) - np.nanmax(s(9), axis=1))
stat('first stamp -> first tile committed', s(2)[:, 0] - t0[:, 0, 0])
for c in (0, ctas // 2):
    print(f'  CTA {c}: per tile [acc owned, slab0, committed | seen, in registers, c0 stored, -, c1 stored | producer, block load, block seen] in ns from the CTA start')
    for i in range(min(tiles, 14)):
        if np.isfinite(T[c, i, 2]):
            print('     ' + ' '.join(f'{(T[c, i, k] - t0[c, 0, 0]) * ns:7.0f}' if np.isfinite(T[c, i, k]) else '      -' for k in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10)))
