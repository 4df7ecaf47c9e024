"""Workloads for ncu.  `step`: one eager flow+warp step between cudaProfilerStart/Stop (launch list);
`kernels`: each hand-written kernel a few times at config-2 size (for --set full captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sd_animation_optical_flow_b200 import ops  # noqa: E402
from sd_animation_optical_flow_b200.engine import RaftEngine  # noqa: E402
from sd_animation_optical_flow_b200.raft import coords_grid  # noqa: E402

dev = torch.device('cuda', 0)
mode = sys.argv[1] if len(sys.argv) > 1 else 'step'
prec = sys.argv[2] if len(sys.argv) > 2 else 'tf32'
g = torch.Generator(device=dev).manual_seed(0)
if mode == 'step':
    cl = len(sys.argv) > 3 and sys.argv[3] == 'cl'
    tcg = len(sys.argv) > 3 and sys.argv[3] == 'tcgru'
    eng = RaftEngine(iters=20, device=dev, corr_precision=prec, channels_last=cl and not tcg, flow_head_scale=0.02, fast_options=dict(tc_gru=tcg))
    a = torch.randint(0, 256, (1, 768, 512, 3), dtype=torch.uint8, device=dev)
    b = a.roll(3, 1)
    sty = torch.randint(0, 256, (1, 768, 512, 3), dtype=torch.uint8, device=dev)
    for _ in range(3):
        ops.warp(sty, eng.estimate_flow(a, b), 'cv2_cubic', -1.0)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    ops.warp(sty, eng.estimate_flow(a, b), 'cv2_cubic', -1.0)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
elif mode == 'gru':
    # the tcgen05 GRU kernels at the batch-1 size
    B, hh, ww = 1, 96, 64
    rnd = lambda *s: torch.randn(s, generator=g, device=dev)
    Hs = torch.tanh(rnd(B, hh, ww, 128))
    hx16 = torch.cat([Hs, torch.relu(rnd(B, hh, ww, 126)), rnd(B, hh, ww, 2)], -1).half().contiguous()
    zrmap, qmap = rnd(B, hh, ww, 256), rnd(B, hh, ww, 128)
    Z, QX = torch.empty_like(Hs), torch.empty_like(Hs)
    RH16 = torch.empty((B, hh, ww, 128), dtype=torch.float16, device=dev)
    wz, wq = ops.gru_weights16(rnd(384, 256, 1, 5) * 0.03), ops.gru_weights16(rnd(128, 128, 1, 5) * 0.05)
    for _ in range(3):
        ops.gru_zr_tc(hx16, wz, zrmap, Hs, True, Z, RH16, QX)
        ops.gru_q_tc(RH16, wq, qmap, QX, Z, True, Hs, hx16)
    torch.cuda.synchronize()
elif mode == 'corr16':
    # the round-2 product path: fp16 operands (auto-ranged), fp16-stored pyramid, lookup reading it
    hh, ww = (90, 160) if prec == 'L' else (96, 64)
    f1 = torch.randn((1, hh, ww, 256), generator=g, device=dev)
    f2 = torch.randn((1, hh, ww, 256), generator=g, device=dev)
    coords_nhwc = (coords_grid(1, hh, ww, dev) + 2 * torch.randn((1, 2, hh, ww), generator=g, device=dev)).permute(0, 2, 3, 1).contiguous()
    look_nhwc = torch.empty((1, hh, ww, 324), device=dev)
    for _ in range(3):
        pyr = ops.corr_volume_pyramid(f1, f2, 4, 'fp16', 'fp16')
        ops.corr_lookup_nhwc(pyr, coords_nhwc, 4, look_nhwc)
    torch.cuda.synchronize()
else:
    f1 = torch.randn((1, 96, 64, 256), generator=g, device=dev)
    f2 = torch.randn((1, 96, 64, 256), generator=g, device=dev)
    coords = coords_grid(1, 96, 64, dev) + 2 * torch.randn((1, 2, 96, 64), generator=g, device=dev)
    coords_nhwc = coords.permute(0, 2, 3, 1).contiguous()
    look_nhwc = torch.empty((1, 96, 64, 324), device=dev)
    src = torch.randint(0, 256, (32, 768, 512, 3), dtype=torch.uint8, device=dev)
    # the flows bench.py uses: translation + low-frequency deformation
    flow = (torch.nn.functional.interpolate(torch.randn((32, 2, 12, 8), generator=g, device=dev) * 4, scale_factor=64, mode='bicubic',
                                            align_corners=False)
            + 6 * torch.randn((32, 2, 1, 1), generator=g, device=dev)).permute(0, 2, 3, 1).contiguous()
    wm = torch.randn((32, 2, 768, 512), generator=g, device=dev) * 3
    act = torch.randn((2, 64, 384, 256), generator=g, device=dev).contiguous(memory_format=torch.channels_last)   # fnet layer-1 size
    lowflow = torch.randn((1, 96, 64, 2), generator=g, device=dev)
    w7 = torch.randn((7, 7, 2, 128), generator=g, device=dev)
    b7 = torch.randn((128,), generator=g, device=dev)
    x256 = torch.relu(torch.randn((1, 96, 64, 256), generator=g, device=dev))
    w2 = torch.randn((3, 3, 2, 256), generator=g, device=dev) * 0.05
    c1 = coords_nhwc.clone()
    fl = torch.empty((1, 96, 64, 2), device=dev)
    mask = (torch.rand((32, 768, 512), generator=g, device=dev) < 0.3).to(torch.uint8) * 255
    # the fp16 update-loop kernels (csrc/raft_glue16.cu, fp16 lookup output) at the batch-1 size
    hf = lambda *sh: (torch.randn(sh, generator=g, device=dev) * 0.5).half()
    corr16 = torch.empty((1, 96, 64, 328), dtype=torch.float16, device=dev)
    f1_16 = torch.empty((1, 96, 64, 128), dtype=torch.float16, device=dev)
    mc16, mf16 = hf(1, 96, 64, 128), hf(1, 96, 64, 128)
    hx16 = torch.empty((1, 96, 64, 256), dtype=torch.float16, device=dev)
    zr16, q16 = hf(1, 96, 64, 384), hf(1, 96, 64, 128)
    zrmap, qmap = torch.randn((1, 96, 64, 256), generator=g, device=dev), torch.randn((1, 96, 64, 128), generator=g, device=dev)
    hid = torch.tanh(torch.randn((1, 96, 64, 128), generator=g, device=dev))
    rh16 = torch.empty((1, 96, 64, 128), dtype=torch.float16, device=dev)
    h16 = torch.empty((1, 96, 64, 128), dtype=torch.float16, device=dev)
    x16 = torch.relu(hf(1, 96, 64, 256))
    bias128 = torch.randn((128,), generator=g, device=dev)
    scratch = torch.empty((96 * 64 * 18,), device=dev)
    i2c16 = torch.empty((1, 96, 64, 200), dtype=torch.float16, device=dev)
    for _ in range(3):
        pyr = ops.corr_volume_pyramid(f1, f2, 4, prec, 'fp16' if prec in ('fp16', 'bf16') else 'fp32')
        ops.corr_lookup(pyr, coords, 4)
        ops.corr_lookup_nhwc(pyr, coords_nhwc, 4, look_nhwc)
        ops.warp(src, flow)
        ops.warp_mask_composite(src[:1], src, flow, wm, 0.95, 7)
        ops.instnorm_nhwc(act, torch.zeros((2 * 64 * 2,), dtype=torch.float64, device=dev))
        ops.conv7x7_c2_relu(lowflow, w7, b7)
        ops.flowhead2_update(x256, w2, (0.1, 0.2), c1, fl, None, 0, None, 0)
        ops.mask_blur_composite(mask, src, src.flip(0), 4.0)
        ops.instnorm_nhwc(act.half().contiguous(memory_format=torch.channels_last), torch.zeros((2 * 64 * 2,), dtype=torch.float64, device=dev))
        if pyr.elem_bytes == 2:
            ops.corr_lookup_nhwc_h(pyr, coords_nhwc, corr16)
        ops.conv7x7_c2_relu_h(lowflow, w7, b7, f1_16)
        ops.motion_tail16_h(mc16, mf16, bias128, lowflow, hx16)
        ops.gru_rh_h(zr16, zrmap, hid, rh16)
        ops.gru_update_h(zr16, zrmap, q16, qmap, hid, hx16, h16)
        ops.flowhead2_update_h(x16, w2, (0.1, 0.2), c1, fl, scratch)
        ops.flow_im2col7_h(c1, scratch, (0.1, 0.2), i2c16)
    torch.cuda.synchronize()
