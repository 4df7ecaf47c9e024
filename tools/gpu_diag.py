"""First-contact diagnostics on the GPU box: runs every kernel family once, prints error statistics
and timings instead of asserting, and keeps going after a failure.  Output -> gpurun_out/diag.txt."""
from __future__ import annotations

import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
LOG = open(os.path.join(ROOT, 'gpurun_out', 'diag.txt'), 'w')


def say(*a):
    msg = ' '.join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + '\n')
    LOG.flush()


def section(fn):
    say(f'--- {fn.__name__}')
    try:
        fn()
    except Exception:
        say('EXCEPTION', traceback.format_exc())


import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import corr_oracle as co  # noqa: E402
from oracle import warp_oracle as wo  # noqa: E402
from sd_animation_optical_flow_b200 import build, ops  # noqa: E402

build.build()
dev = torch.device('cuda', 0)
say(torch.cuda.get_device_name(0), torch.version.cuda, 'SMs', torch.cuda.get_device_properties(0).multi_processor_count)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def timeit(fn, n=20):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3  # us


def corr_small():
    rs = np.random.RandomState(0)
    for shape in ((1, 32, 8, 32), (1, 256, 16, 40), (2, 64, 17, 23)):
        f1 = rs.standard_normal(shape).astype(np.float32)
        f2 = rs.standard_normal(shape).astype(np.float32)
        ref = co.corr_pyramid(f1, f2, 4)
        for prec in ('fp32', 'fp16', 'tf32', '3xtf32', 'bf16'):
            try:
                pyr = ops.corr_volume_pyramid(t(f1).permute(0, 2, 3, 1).contiguous(), t(f2).permute(0, 2, 3, 1).contiguous(), 4, prec)
                torch.cuda.synchronize()
                errs = []
                for l in range(4):
                    got = pyr.level(l)[:, 0].cpu().numpy()
                    errs.append(float(np.abs(got - ref[l]).max()) if got.size else 0.0)
                say(shape, prec, 'max|ref|', float(np.abs(ref[0]).max()), 'level errs', ['%.2e' % e for e in errs])
                if prec != 'fp32' and errs[0] > 0.1:
                    got = pyr.level(0)[:, 0].cpu().numpy()
                    d = np.abs(got - ref[0])
                    bad = np.argwhere(d > 0.1)
                    say('   bad count', len(bad), 'of', d.size, 'first', bad[:8].tolist())
                    say('   got[0,0,:8]', got[0, 0, :8], 'ref', ref[0][0, 0, :8])
                    say('   rows with errors (mod 32 histogram)', np.bincount(bad[:, 0] % 32, minlength=32).tolist())
                    say('   cols with errors (x histogram)', np.bincount(bad[:, 2], minlength=got.shape[2]).tolist())
                    say('   rows y histogram', np.bincount(bad[:, 1], minlength=got.shape[1]).tolist())
            except Exception:
                say(shape, prec, 'EXCEPTION', traceback.format_exc())


def corr_perf():
    g = torch.Generator(device=dev).manual_seed(0)
    for (h, w) in ((96, 64), (90, 160)):
        f1 = torch.randn((1, h, w, 256), generator=g, device=dev)
        f2 = torch.randn((1, h, w, 256), generator=g, device=dev)
        n = h * w
        for prec in ('fp16', 'bf16', 'tf32', '3xtf32', 'fp32'):
            try:
                us = timeit(lambda: ops.corr_volume_pyramid(f1, f2, 4, prec), 10 if prec != 'fp32' else 3)
                lay = ops.corr_volume_pyramid(f1, f2, 4, prec).layout
                out_b = 4 * n * sum(lay.h[l] * lay.w[l] for l in range(4))
                say(f'{h}x{w} {prec}: {us:.1f} us  {2.0 * n * n * 256 / us / 1e6:.1f} TFLOP/s  store {out_b / us / 1e3:.0f} GB/s')
            except Exception:
                say(h, w, prec, 'EXCEPTION', traceback.format_exc())
        # reference path for scale: torch matmul + scale + 3 pools (what CorrBlock does)
        a = f1.reshape(1, n, 256)
        b = f2.reshape(1, n, 256)

        def ref_path():
            c = torch.matmul(a, b.transpose(1, 2)) / 16.0
            c = c.reshape(n, 1, h, w)
            for _ in range(3):
                c = torch.nn.functional.avg_pool2d(c, 2, stride=2)
            return c
        say(f'{h}x{w} torch CorrBlock-equivalent build (fp32 cuBLAS + scale + 3 avg_pool2d): {timeit(ref_path, 5):.1f} us')


def lookup_perf():
    g = torch.Generator(device=dev).manual_seed(0)
    f1 = torch.randn((1, 96, 64, 256), generator=g, device=dev)
    f2 = torch.randn((1, 96, 64, 256), generator=g, device=dev)
    pyr = ops.corr_volume_pyramid(f1, f2, 4, 'tf32')
    from sd_animation_optical_flow_b200.raft import coords_grid
    coords = coords_grid(1, 96, 64, dev) + 2 * torch.randn((1, 2, 96, 64), generator=g, device=dev)
    out = torch.empty((1, 324, 96, 64), device=dev)
    us = timeit(lambda: ops.corr_lookup(pyr, coords, 4, out=out), 50)
    say(f'lookup 96x64: {us:.1f} us  {2896.0 * 6144 / us / 1e3:.0f} GB/s algorithmic')


def warp_perf():
    g = torch.Generator(device=dev).manual_seed(0)
    for B in (1, 32):
        src = torch.randint(0, 256, (B, 768, 512, 3), dtype=torch.uint8, device=dev)
        rough = torch.randn((B, 768, 512, 2), generator=g, device=dev) * 6
        us = timeit(lambda: ops.warp(src, rough), 20)
        say(f'warp cubic u8 B={B} (adversarial per-pixel random flow): {us:.1f} us  {14.0 * B * 768 * 512 / us / 1e3:.0f} GB/s algorithmic')
        flow = torch.nn.functional.interpolate(torch.randn((B, 2, 96, 64), generator=g, device=dev) * 6, scale_factor=8, mode='bilinear',
                                               align_corners=False).permute(0, 2, 3, 1).contiguous()
        us = timeit(lambda: ops.warp(src, flow), 20)
        say(f'warp cubic u8 B={B} (smooth flow, 8x-upsampled field): {us:.1f} us  {14.0 * B * 768 * 512 / us / 1e3:.0f} GB/s algorithmic')
        us = timeit(lambda: ops.warp(src, flow, 'bilinear'), 20)
        say(f'warp bilinear u8 B={B}: {us:.1f} us')
        wm = torch.randn((B, 2, 768, 512), generator=g, device=dev) * 3
        us = timeit(lambda: ops.warp_mask_composite(src[:1], src, flow, wm, 0.95, 7), 20)
        say(f'fused warp+mask+composite B={B}: {us:.1f} us  {26.0 * B * 768 * 512 / us / 1e3:.0f} GB/s algorithmic')


def warp_exact():
    rs = np.random.RandomState(1)
    img = rs.randint(0, 256, (200, 160, 3)).astype(np.uint8)
    flow = (8 * rs.standard_normal((200, 160, 2))).astype(np.float32)
    out = ops.warp(t(img), t(flow)).cpu().numpy()
    ref = wo.warp_frame_pdcnet(img, flow)
    say('cubic u8 mismatching bytes:', int((out != ref).sum()), 'of', ref.size)
    if (out != ref).any():
        bad = np.argwhere(out != ref)
        say('   first', bad[:10].tolist(), 'got', out[tuple(bad[0])], 'ref', ref[tuple(bad[0])])


def raft_perf():
    from sd_animation_optical_flow_b200.engine import RaftEngine
    img = torch.randint(0, 256, (1, 768, 512, 3), dtype=torch.uint8, device=dev)
    for kw in (dict(fast=False), dict(fast=False, use_cuda_graph=True), dict(fast=True, use_cuda_graph=False), dict(fast=True, use_cuda_graph=True)):
        try:
            eng = RaftEngine(iters=20, device=dev, **kw)
            t0 = time.time()
            eng.estimate_flow(img, img.flip(1))
            torch.cuda.synchronize()
            first = time.time() - t0
            us = timeit(lambda: eng.estimate_flow(img, img.flip(1)), 5)
            say(f'RAFT 768x512 iters=20 {kw}: {us / 1e3:.2f} ms/pair (first call {first:.1f} s)')
        except Exception:
            say(kw, 'EXCEPTION', traceback.format_exc())


for fn in (corr_small, warp_exact, corr_perf, lookup_perf, warp_perf, raft_perf):
    section(fn)
say('diag done')
