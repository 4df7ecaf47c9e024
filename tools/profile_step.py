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
    eng = RaftEngine(iters=20, device=dev, corr_precision=prec, channels_last=cl)
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
else:
    f1 = torch.randn((1, 96, 64, 256), generator=g, device=dev)
    f2 = torch.randn((1, 96, 64, 256), generator=g, device=dev)
    coords = coords_grid(1, 96, 64, dev) + 2 * torch.randn((1, 2, 96, 64), generator=g, device=dev)
    src = torch.randint(0, 256, (32, 768, 512, 3), dtype=torch.uint8, device=dev)
    flow = torch.nn.functional.interpolate(torch.randn((32, 2, 96, 64), generator=g, device=dev) * 6, scale_factor=8, mode='bilinear',
                                           align_corners=False).permute(0, 2, 3, 1).contiguous()
    wm = torch.randn((32, 2, 768, 512), generator=g, device=dev) * 3
    for _ in range(3):
        pyr = ops.corr_volume_pyramid(f1, f2, 4, prec)
        ops.corr_lookup(pyr, coords, 4)
        ops.warp(src, flow)
        ops.warp_mask_composite(src[:1], src, flow, wm, 0.95, 7)
    torch.cuda.synchronize()
