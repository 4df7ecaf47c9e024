"""CUDA-event timings of the correlation kernels at config sizes: operand pre-pass, pyramid kernel (fp32 / fp16 storage),
lookup (planar / channels-last, fp32 / fp16 pyramid).   python tools/corr_bench.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sd_animation_optical_flow_b200 import ops  # noqa: E402
from sd_animation_optical_flow_b200.raft import coords_grid  # noqa: E402


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


def main():
    dev = torch.device('cuda', 0)
    for (h, w) in ((96, 64), (90, 160)):
        g = torch.Generator(device=dev).manual_seed(0)
        f1 = torch.randn((1, h, w, 256), generator=g, device=dev)
        f2 = torch.randn((1, h, w, 256), generator=g, device=dev)
        n1 = h * w
        row = {'hw': [h, w]}
        row['prepare_src_us'] = t_us(lambda: ops.CorrSource(f1, 'fp16'))
        row['prepare_tgt_us'] = t_us(lambda: ops.CorrTarget(f2, 4, 'fp16'))
        src, tgt = ops.CorrSource(f1, 'fp16'), ops.CorrTarget(f2, 4, 'fp16')
        for storage in ('fp16', 'fp32'):
            pyr = src.pyramid(tgt, storage)
            row[f'pyramid_kernel_{storage}_us'] = t_us(lambda: src.pyramid(tgt, storage, out=pyr))
            row[f'pyramid_op_{storage}_us'] = t_us(lambda: ops.corr_volume_pyramid(f1, f2, 4, 'fp16', storage), 10)
            lay = pyr.layout
            out_bytes = pyr.elem_bytes * n1 * sum(lay.h[l] * lay.w[l] for l in range(4))
            row[f'pyramid_{storage}_out_MB'] = round(out_bytes / 1e6, 1)
            row[f'pyramid_kernel_{storage}_TBps'] = round((out_bytes + n1 * 256 * 2 * (1 + 85 / 64)) / row[f'pyramid_kernel_{storage}_us'] / 1e6, 3)
            coords = coords_grid(1, h, w, dev) + 2 * torch.randn((1, 2, h, w), generator=g, device=dev)
            cn = coords.permute(0, 2, 3, 1).contiguous()
            o1 = torch.empty((1, 324, h, w), device=dev)
            o2 = torch.empty((1, h, w, 324), device=dev)
            row[f'lookup_planar_{storage}_us'] = t_us(lambda: ops.corr_lookup(pyr, coords, 4, out=o1), 50)
            row[f'lookup_nhwc_{storage}_us'] = t_us(lambda: ops.corr_lookup_nhwc(pyr, cn, 4, o2), 50)
            del pyr
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
