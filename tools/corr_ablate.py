"""Ablation timings of corr_pyramid_resident_kernel via its SDOF_RES_DEBUG switches (1 skip stores, 4 skip MMA, 8 skip the
target-slab TMA loads, 16 skip the TMEM loads): which stage bounds the kernel.   python tools/corr_ablate.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sd_animation_optical_flow_b200 import ops  # noqa: E402


def t_us(fn, n=20):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return round(s.elapsed_time(e) / n * 1e3, 2)


dev = torch.device('cuda', 0)
for (h, w) in ((96, 64), (90, 160)):
    g = torch.Generator(device=dev).manual_seed(0)
    f1 = torch.randn((1, h, w, 256), generator=g, device=dev)
    f2 = torch.randn((1, h, w, 256), generator=g, device=dev)
    src, tgt = ops.CorrSource(f1, 'fp16'), ops.CorrTarget(f2, 4, 'fp16')
    for storage in ('fp16', 'fp32'):
        pyr = src.pyramid(tgt, storage)
        row = {'hw': [h, w], 'storage': storage}
        for name, dbg in (('full', 0), ('no_stores', 1), ('no_mma', 4), ('no_loads', 8), ('no_tmem_ld', 16), ('no_stores_no_mma', 5),
                          ('no_stores_no_loads', 9), ('only_stores', 4 | 8 | 16), ('nothing', 1 | 4 | 8 | 16)):
            os.environ['SDOF_RES_DEBUG'] = str(dbg)
            row[name] = t_us(lambda: src.pyramid(tgt, storage, out=pyr))
        os.environ['SDOF_RES_DEBUG'] = '0'
        print(json.dumps(row), flush=True)
