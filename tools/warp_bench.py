"""Times the warp-family kernels at config-2/3 sizes (CUDA events, back-to-back launches) and prints their share of
the measured HBM peak.  Usage: python tools/warp_bench.py  (GPU box)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sd_animation_optical_flow_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
HBM = peaks['hbm_gbs']


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e-3


def smooth_flow(B, H, W, amp):
    return torch.nn.functional.interpolate(torch.randn((B, 2, H // 8, W // 8), generator=g, device=dev) * amp, scale_factor=8,
                                           mode='bilinear', align_corners=False).permute(0, 2, 3, 1).contiguous()


for (B, H, W) in ((32, 768, 512), (1, 768, 512), (16, 720, 1280)):
    src = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    wm = torch.randn((B, 2, H, W), generator=g, device=dev) * 3
    for name, flow in (('smooth6', smooth_flow(B, H, W, 6.0)), ('smooth1', smooth_flow(B, H, W, 1.0)),
                       ('zero', torch.zeros((B, H, W, 2), device=dev)),
                       ('noise12', 12 * torch.randn((B, H, W, 2), generator=g, device=dev))):
        t = timeit(lambda: ops.warp(src, flow))
        gbs = 14.0 * B * H * W / t / 1e9
        print(f'warp_cubic_u8c3   B={B:2d} {H}x{W} flow={name:8s} {t * 1e6:8.1f} us  {gbs:7.1f} GB/s  {gbs / HBM * 100:5.1f}% of HBM peak')
        if name in ('smooth6', 'noise12'):
            t = timeit(lambda: ops.warp_mask_composite(src[:1], src, flow, wm, 0.95, 7))
            gbs = 26.0 * B * H * W / t / 1e9
            print(f'warp_mask_composite B={B:2d} {H}x{W} flow={name:8s} {t * 1e6:8.1f} us  {gbs:7.1f} GB/s  {gbs / HBM * 100:5.1f}% of HBM peak')
    t = timeit(lambda: ops.warp(src, flow, 'bilinear'))
    print(f'warp_bilinear_u8  B={B:2d} {H}x{W} {t * 1e6:8.1f} us')

# after-the-path step: mask blur + composite (11 B/pixel) and the latent-mask resize
B, H, W = 32, 768, 512
src = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
ref = src.flip(0).contiguous()
mask = ((torch.rand((B, H // 16, W // 16), generator=g, device=dev) < 0.3).to(torch.uint8) * 255).repeat_interleave(16, 1).repeat_interleave(16, 2).contiguous()
for blur in (4.0, 12.0):
    t = timeit(lambda: ops.mask_blur_composite(mask, src, ref, blur))
    gbs = 11.0 * B * H * W / t / 1e9
    print(f'blur_composite    B={B:2d} {H}x{W} mask_blur={blur:4.1f} {t * 1e6:8.1f} us  {gbs:7.1f} GB/s  {gbs / HBM * 100:5.1f}% of HBM peak')
t = timeit(lambda: ops.resize_bicubic_u8(mask, H // 8, W // 8, want_latmask=True), 5)
print(f'resize_bicubic_u8 B={B:2d} {H}x{W} -> {H // 8}x{W // 8} {t * 1e6:8.1f} us (includes the host-built coefficient tables and one stream sync)')
frame = src[0].contiguous()
from sd_animation_optical_flow_b200 import ofgen  # noqa: E402
t = timeit(lambda: ofgen.detect_edges_device(frame), 5)
print(f'detect_edges      768x512 {t * 1e6:8.1f} us per frame (V + histogram + thresholds + Canny + hysteresis relaunches + dilation)')
