"""In-graph timings of the tcgen05 GRU kernels (csrc/conv_tc.cu) and of the cuDNN + glue sequence they replace, at the
batch-1 sizes of bench.py.   python tools/gru_bench.py [H W]"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sd_animation_optical_flow_b200 import ops  # noqa: E402

CL = torch.channels_last


def time_graphed(fn, reps=20, rounds=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    best = 1e9
    for _ in range(rounds):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / reps * 1e3)
    return round(best, 2)


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (768, 512)
    B, h, w = (int(sys.argv[3]) if len(sys.argv) > 3 else 1), H // 8, W // 8
    dev = torch.device('cuda', 0)
    torch.backends.cudnn.benchmark = True
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.randn(s, generator=g, device=dev)
    row = {'hw': [h, w], 'B': B}
    for horizontal in (True, False):
        tag = '1x5' if horizontal else '5x1'
        ks = (1, 5) if horizontal else (5, 1)
        pad = (0, 2) if horizontal else (2, 0)
        w_zr, w_q = rnd(384, 256, *ks) * 0.03, rnd(128, 128, *ks) * 0.05
        Hs = torch.tanh(rnd(B, h, w, 128))
        hx16 = torch.cat([Hs, torch.relu(rnd(B, h, w, 126)), rnd(B, h, w, 2)], -1).half().contiguous()
        zrmap, qmap = rnd(B, h, w, 256), rnd(B, h, w, 128)
        Z, QX = torch.empty_like(Hs), torch.empty_like(Hs)
        RH16 = torch.empty((B, h, w, 128), dtype=torch.float16, device=dev)
        wz, wq = ops.gru_weights16(w_zr), ops.gru_weights16(w_q)
        row[f'gru_zr_tc_{tag}_us'] = time_graphed(lambda: ops.gru_zr_tc(hx16, wz, zrmap, Hs, horizontal, Z, RH16, QX))
        row[f'gru_q_tc_{tag}_us'] = time_graphed(lambda: ops.gru_q_tc(RH16, wq, qmap, QX, Z, horizontal, Hs, hx16))
        # what they replace: cuDNN TF32 convolutions + glue kernels
        HX = hx16.float()
        RH = torch.empty_like(Hs)
        wzc, wqc = w_zr.contiguous(memory_format=CL), w_q.contiguous(memory_format=CL)

        def conv(x, wt):
            y = F.conv2d(x.permute(0, 3, 1, 2), wt, None, padding=pad)
            return (y if y.is_contiguous(memory_format=CL) else y.contiguous(memory_format=CL)).permute(0, 2, 3, 1)
        row[f'cudnn_zr_{tag}_us'] = time_graphed(lambda: conv(HX, wzc))
        row[f'cudnn_q_{tag}_us'] = time_graphed(lambda: conv(RH, wqc))
        zr = conv(HX, wzc)
        qq = conv(RH, wqc)
        row[f'glue_gru_rh_{tag}_us'] = time_graphed(lambda: ops.gru_rh(zr, Hs, RH, bias_zr=zrmap))
        row[f'glue_gru_update_{tag}_us'] = time_graphed(lambda: ops.gru_update(zr, qq, Hs, HX, bias_zr=zrmap, bias_q=qmap))
    mc, mf, bias, flow = rnd(B, h, w, 128), rnd(B, h, w, 128), rnd(128), rnd(B, h, w, 2)
    row['motion_tail16_us'] = time_graphed(lambda: ops.motion_tail16(mc, mf, bias, flow, hx16))
    print(json.dumps(row))


if __name__ == '__main__':
    main()
