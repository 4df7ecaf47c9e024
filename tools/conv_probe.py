"""Times every convolution shape of the RAFT update block (and the encoders' bulk shapes) through cuDNN in
TF32 / fp16 / bf16, channels-last, at the batch-1 sizes bench.py uses.  Each shape is replayed 20x inside a CUDA
graph so the launch overhead of eager PyTorch is not part of the number.

    python tools/conv_probe.py [H W]        # default 768 512
"""
import json
import sys

import torch
import torch.nn.functional as F

CL = torch.channels_last


def time_graph(fn, reps=20, rounds=5):
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
    return best


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (768, 512)
    h, w = H // 8, W // 8
    dev = torch.device('cuda', 0)
    torch.backends.cudnn.benchmark = True
    # (name, Cin, Cout, kh, kw, spatial h, spatial w, batch, stride)
    shapes = [
        ('convc1 1x1 324->256', 324, 256, 1, 1, h, w, 1, 1),
        ('convc2 3x3 256->192', 256, 192, 3, 3, h, w, 1, 1),
        ('convf2 3x3 128->64', 128, 64, 3, 3, h, w, 1, 1),
        ('conv_cor 3x3 192->128', 192, 128, 3, 3, h, w, 1, 1),
        ('conv_flo 3x3 64->128', 64, 128, 3, 3, h, w, 1, 1),
        ('zr 1x5 256->384', 256, 384, 1, 5, h, w, 1, 1),
        ('zr 5x1 256->384', 256, 384, 5, 1, h, w, 1, 1),
        ('q 1x5 128->128', 128, 128, 1, 5, h, w, 1, 1),
        ('q 5x1 128->128', 128, 128, 5, 1, h, w, 1, 1),
        ('fh1 3x3 128->256', 128, 256, 3, 3, h, w, 1, 1),
        ('mask0 3x3 128->256', 128, 256, 3, 3, h, w, 1, 1),
        ('mask2 1x1 256->576', 256, 576, 1, 1, h, w, 1, 1),
        ('enc stem 7x7 4->64 s2', 4, 64, 7, 7, H, W, 2, 2),
        ('enc l1 3x3 64->64', 64, 64, 3, 3, H // 2, W // 2, 2, 1),
        ('enc l2 3x3 64->96 s2', 64, 96, 3, 3, H // 2, W // 2, 2, 2),
        ('enc l2 3x3 96->96', 96, 96, 3, 3, H // 4, W // 4, 2, 1),
        ('enc l3 3x3 96->128 s2', 96, 128, 3, 3, H // 4, W // 4, 2, 2),
        ('enc l3 3x3 128->128', 128, 128, 3, 3, H // 8, W // 8, 2, 1),
        ('enc out 1x1 128->256', 128, 256, 1, 1, H // 8, W // 8, 2, 1),
    ]
    rows = []
    for name, ci, co, kh, kw, sh, sw, b, st in shapes:
        row = {'shape': name, 'gflop': 2.0 * b * (sh // st) * (sw // st) * ci * co * kh * kw / 1e9}
        for tag, dt, tf32 in (('fp32', torch.float32, False), ('tf32', torch.float32, True), ('fp16', torch.float16, True),
                              ('bf16', torch.bfloat16, True)):
            if tag == 'fp32' and row['gflop'] > 8:
                continue
            torch.backends.cudnn.allow_tf32 = tf32
            x = torch.randn((b, ci, sh, sw), device=dev, dtype=dt).contiguous(memory_format=CL)
            wt = (torch.randn((co, ci, kh, kw), device=dev, dtype=dt) * 0.05).contiguous(memory_format=CL)
            pad = (kh // 2, kw // 2)
            try:
                row[tag + '_us'] = round(time_graph(lambda: F.conv2d(x, wt, None, stride=st, padding=pad)), 2)
            except Exception as e:  # noqa: BLE001
                row[tag + '_us'] = f'error {type(e).__name__}'
        rows.append(row)
        print(json.dumps(row), flush=True)
    tot = {k: sum(r[k] for r in rows[:12] if isinstance(r.get(k), float)) for k in ('tf32_us', 'fp16_us', 'bf16_us')}
    print(json.dumps({'update_block_sum': tot, 'H': H, 'W': W}))


if __name__ == '__main__':
    main()
