"""Benchmark of the flow -> warp hot path (BASELINE.json metric: frame-pairs/sec (flow+warp) at 512x768;
corr-volume tensor-pipe % of peak).

    python bench.py --gpus 1 --steps 20 --warmup 3                      # our arm
    python bench.py --impl reference --gpus 1 --steps 5 --warmup 1      # CPU reference arm
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W       # N > 1, one rank per GPU

A "step" is one pass of the hot path over one synthetic 768x512 frame pair per rank (config 2:
RAFT all-pairs correlation + warp, random-init weights, iters=20): estimate_flow + cubic warp of the
stylised previous frame.  Pairs are independent, so ranks hold different pairs and there is no
collective on the data path (weak scaling); `value` = pairs all ranks processed / max-over-ranks time.

Prints ONE JSON line (rank 0).  Extra keys: roofline (dominant hand-written kernel), roofline_extra,
cpu_baseline, e2e, clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 768, 512          # tensors are H=768, W=512 (cv2.resize(frame,(512,768)), ofgen_pixel_inpaint.py:324)
ITERS = 20               # ofgen.py:77
METRIC = 'frame-pairs/sec (flow+warp) at 512x768'
UNIT = 'frame-pairs/s'


def synthetic_pair(seed: int):
    """Gaussian-blurred noise texture; frame 2 = frame 1 shifted by (+4,-3) px."""
    from tests import golden_inputs as gi
    f1, f2 = gi.shifted_pair(H, W, 1000 + seed)
    stylised = gi.texture(H, W, 2000 + seed)
    return f1, f2, stylised


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi sampling DURING the timed regions (B200_PROFILING.md recipe).  Started before the warm-up steps so it is
    already polling when the timed region begins; 100 ms period (a 50 ms period perturbed the host-timed e2e loop on one
    box through driver-lock contention)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        os.unlink(self.path)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def cpu_baseline_leg(budget_s: float):
    from oracle import farneback_baseline as fb
    fb.use_all_host_threads()
    f1, f2, sty = synthetic_pair(0)
    bgr = lambda a: a[:, :, ::-1].copy()
    rate, n, times = fb.time_pairs(bgr(f1), bgr(f2), bgr(sty), budget_s=budget_s)
    info = fb.host_info()
    return {'value': rate, 'unit': UNIT, 'cores': info['cores'], 'kind': 'port',
            'sample': f'{n} x one 768x512 pair, cv2.calcOpticalFlowFarneback(0.5,5,15,3,5,1.2,0) + torch grid_sample warp; '
                      f'best {min(times) * 1e3:.1f} ms, median {statistics.median(times) * 1e3:.1f} ms',
            'cpu_model': info['cpu_model'], 'cv2_threads': info['cv2_threads'], 'torch_threads': info['torch_threads']}


def run_reference(args):
    """The reference's CPU path for this metric (north star: Farneback + grid_sample on the host cores),
    all host threads; rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    from oracle import farneback_baseline as fb
    fb.use_all_host_threads()
    f1, f2, sty = synthetic_pair(0)
    bgr = lambda a: a[:, :, ::-1].copy()
    a, b, c = bgr(f1), bgr(f2), bgr(sty)
    for _ in range(max(args.warmup, 1)):
        fb.flow_and_warp(a, b, c)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fb.flow_and_warp(a, b, c)
    dt = time.perf_counter() - t0
    info = fb.host_info()
    value = args.steps / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'configs[1]: single 512x768 frame pair, CPU Farneback flow + grid_sample warp', 'H': H, 'W': W},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': info['cores'], 'kind': 'port',
                             'sample': f'{args.steps} x one 768x512 pair per step', 'cpu_model': info['cpu_model'],
                             'cv2_threads': info['cv2_threads'], 'torch_threads': info['torch_threads']},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def time_op(fn, n: int, torch):
    """Average device time of fn() over n calls, CUDA events on the current stream."""
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e-3


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from sd_animation_optical_flow_b200 import _capi, ofgen, ops
    from sd_animation_optical_flow_b200 import build as _build
    from sd_animation_optical_flow_b200.engine import RaftEngine

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    _capi.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    f1, f2, sty = synthetic_pair(rank)
    d1 = torch.from_numpy(f1).to(dev)[None]
    d2 = torch.from_numpy(f2).to(dev)[None]
    dsty = torch.from_numpy(sty).to(dev)[None]
    eng = RaftEngine(checkpoint=None, iters=ITERS, corr_precision=args.corr_precision, use_cuda_graph=not args.no_graph,
                     mixed_precision=args.mixed_precision, channels_last=args.channels_last, device=dev, seed=0,
                     cudnn_benchmark=not args.no_cudnn_benchmark,
                     fast_options=dict(side_streams=not args.no_side_streams, own_convf1=not args.cudnn_convf1, own_fh2=not args.cudnn_fh2))

    def step():
        flow = eng.estimate_flow(d1, d2)                 # [1,768,512,2]
        return ops.warp(dsty, flow, 'cv2_cubic', -1.0)   # ofgen.warp_frame: previous stylised frame at x - flow

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    # kernels of THIS library per step, counted on one eager (non-graph) step: graph replays re-run
    # exactly these launches without going through the host-side counter
    graphed = eng.use_cuda_graph
    eng.use_cuda_graph = False
    c0 = _capi.launch_count()
    step()
    launches_per_step = _capi.launch_count() - c0
    eng.use_cuda_graph = graphed
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s.record()
    for _ in range(args.steps):
        out = step()
    e.record()
    barrier()
    dt = s.elapsed_time(e) * 1e-3
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if (rank == 0 and args.quick) else None
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    value = world * args.steps / dt
    if args.quick:
        if rank == 0:
            sampler_stats = clocks
            print(json.dumps({'quick': True, 'value': value, 'ms_per_step': dt / args.steps * 1e3, 'launches_per_step': launches_per_step,
                              'clocks': sampler_stats, 'argv': sys.argv[1:]}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: the reference-facing numpy API with HOST buffers (H2D / D2H inside the timed region)
    algo = ofgen.RAFT_2.__new__(ofgen.RAFT_2)
    algo.engine = eng
    # host buffers live in pinned memory (numpy views of pinned tensors), as the contract asks
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    bgr1, bgr2, bgrs = pin(f1[:, :, ::-1]), pin(f2[:, :, ::-1]), pin(sty[:, :, ::-1])

    def e2e_step():
        flow = algo.calc(bgr1, bgr2)                      # H2D 2 frames, D2H flow
        return ofgen.warp_frame(bgrs, flow)               # H2D frame + flow, D2H warped frame

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        warped_np = e2e_step()
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if rank == 0:
        clocks = sampler.stop()   # sampled across both timed regions (device-resident steps, then the e2e steps)
    if world > 1:
        t = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    frame_b, flow_b = H * W * 3, H * W * 2 * 4
    e2e = {'value': world * args.steps / e2e_dt, 'unit': UNIT, 'h2d_bytes_per_step': 3 * frame_b + flow_b,
           'd2h_bytes_per_step': flow_b + frame_b, 'api': 'ofgen.RAFT_2.calc(np,np) + ofgen.warp_frame(np,np)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the hand-written kernels, timed alone with CUDA events on the launching stream
    peaks = load_peaks()
    n1 = (H // 8) * (W // 8)
    C = 256
    g = torch.Generator(device=dev).manual_seed(0)
    fm1 = torch.randn((1, H // 8, W // 8, C), generator=g, device=dev)
    fm2 = torch.randn((1, H // 8, W // 8, C), generator=g, device=dev)
    pyr_holder = {}

    def corr_op():
        pyr_holder['p'] = ops.corr_volume_pyramid(fm1, fm2, 4, args.corr_precision)

    t_corr_op = time_op(corr_op, 20, torch)   # pre-pass + tensor-core kernel (what one pair pays)
    pyr = pyr_holder['p']
    lay = pyr.layout
    if args.corr_precision in ('fp16', 'bf16'):
        # the dominant kernel alone: operands prepared once (as for pairs sharing a key frame), output preallocated
        operands = ops.CorrOperands(1, H // 8, W // 8, H // 8, W // 8, C, 4, args.corr_precision, dev).prepare(fm1, fm2)
        t_corr = time_op(lambda: operands.pyramid(out=pyr), 20, torch)
        corr_kernel = 'corr_pyramid_resident_kernel'
    else:
        t_corr = t_corr_op
        corr_kernel = 'corr_volume_tc_kernel + operand pre-pass'
    out_bytes = 4 * n1 * sum(lay.h[l] * lay.w[l] for l in range(4))
    in_bytes = 2 * n1 * C * 4
    corr_flops = 2.0 * n1 * n1 * C
    from sd_animation_optical_flow_b200.raft import coords_grid
    coords = coords_grid(1, H // 8, W // 8, dev) + 2 * torch.randn((1, 2, H // 8, W // 8), generator=g, device=dev)
    look_out = torch.empty((1, 324, H // 8, W // 8), device=dev)
    t_look = time_op(lambda: ops.corr_lookup(pyr, coords, 4, out=look_out), 50, torch)
    coords_nhwc = coords.permute(0, 2, 3, 1).contiguous()
    look_nhwc = torch.empty((1, H // 8, W // 8, 324), device=dev)
    t_look_nhwc = time_op(lambda: ops.corr_lookup_nhwc(pyr, coords_nhwc, 4, look_nhwc), 50, torch)
    # flows as the path produces them: smooth fields (camera / object motion): per frame a random translation of a few
    # pixels plus a low-frequency deformation (1/64-resolution Gaussian field of 4 px, bicubic-upsampled: |grad| ~ 0.1)
    flow32 = (torch.nn.functional.interpolate(torch.randn((32, 2, H // 64, W // 64), generator=g, device=dev) * 4, scale_factor=64,
                                              mode='bicubic', align_corners=False)
              + 6 * torch.randn((32, 2, 1, 1), generator=g, device=dev)).permute(0, 2, 3, 1).contiguous()
    src32 = torch.randint(0, 256, (32, H, W, 3), dtype=torch.uint8, device=dev)
    t_warp = time_op(lambda: ops.warp(src32, flow32), 20, torch)
    wm32 = torch.randn((32, 2, H, W), generator=g, device=dev) * 3
    t_fused = time_op(lambda: ops.warp_mask_composite(src32[:1], src32, flow32, wm32, 0.95, 7), 20, torch)

    hbm = peaks['hbm_gbs']
    corr_gbs = (in_bytes + out_bytes) / t_corr / 1e9
    roofline = {'kernel': corr_kernel + ', 1 pair, N=6144, C=256, 4 levels', 'us_per_op_with_prepass': t_corr_op * 1e6, 'bound': 'hbm',
                'achieved': corr_gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': corr_gbs / hbm,
                'traffic': 152921088 if args.corr_precision == 'fp16' else None,
                'traffic_note': 'dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r1b_kernels_ncu_summary.txt: 8.84 MB read + 144.08 MB written); below the algorithmic bytes because part of the 200 MB pyramid is still dirty in the 126 MB L2 when the kernel ends',
                'peak_source': peaks['source'], 'us_per_launch': t_corr * 1e6,
                'algorithmic_bytes': in_bytes + out_bytes}
    tf = corr_flops / t_corr / 1e12
    extra = [
        {'kernel': corr_kernel, 'bound': 'tensor', 'achieved': tf, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
         'frac': tf / peaks['bf16_tflops'], 'note': f'{args.corr_precision} MMA (2*N^2*C flops of level 0) vs measured bf16 burst peak; the kernel is HBM-store-bound'},
        {'kernel': 'corr_lookup_kernel<4,4,32> (planar output, the corr_fn protocol)', 'bound': 'hbm', 'achieved': 2896.0 * n1 / t_look / 1e9,
         'peak': hbm, 'unit': 'GB/s', 'frac': 2896.0 * n1 / t_look / 1e9 / hbm, 'us_per_launch': t_look * 1e6},
        {'kernel': 'corr_lookup_kernel<4,4,8> (channels-last output, the one in the step)', 'bound': 'hbm',
         'achieved': 2896.0 * n1 / t_look_nhwc / 1e9, 'peak': hbm, 'unit': 'GB/s', 'frac': 2896.0 * n1 / t_look_nhwc / 1e9 / hbm,
         'us_per_launch': t_look_nhwc * 1e6},
        {'kernel': 'warp_cubic_u8c3_tiled_kernel, 32 frames, smooth flow (translation + low-frequency deformation)', 'bound': 'hbm', 'achieved': 14.0 * 32 * H * W / t_warp / 1e9, 'peak': hbm,
         'unit': 'GB/s', 'frac': 14.0 * 32 * H * W / t_warp / 1e9 / hbm, 'us_per_launch': t_warp * 1e6},
        {'kernel': 'warp_mask_composite_kernel, 32 frames, same flows, N(0,3) logits, thres 0.95, 7x7 ellipse', 'bound': 'hbm', 'achieved': 26.0 * 32 * H * W / t_fused / 1e9, 'peak': hbm,
         'unit': 'GB/s', 'frac': 26.0 * 32 * H * W / t_fused / 1e9 / hbm, 'us_per_launch': t_fused * 1e6},
    ]

    # ---- the same path with 8 pairs per call (how configs[2..4] feed it: PDCNetAux batches 16 pairs,
    # ofgen_keyframe_inpaint.py:550,585-600); reported beside the single-pair headline, not instead of it
    P = 8
    bd1, bd2, bsty = d1.repeat(P, 1, 1, 1), d2.repeat(P, 1, 1, 1).roll(1, 0), dsty.repeat(P, 1, 1, 1)

    def batched_step():
        return ops.warp(bsty, eng.estimate_flow(bd1, bd2), 'cv2_cubic', -1.0)

    t_b = time_op(batched_step, 5, torch)
    batched = {'pairs_per_step': P, 'ms_per_step': t_b * 1e3, 'value': P / t_b, 'unit': UNIT, 'n_gpus': 1,
               'note': 'device-resident, one GPU (rank 0), CUDA graph' if not args.no_graph else 'device-resident, one GPU (rank 0)'}

    # ---- configs[2] in miniature: a 32-frame clip against its key frame, all on the device: flow + confidence (RAFT has
    # no confidence head: forward-backward consistency, i.e. two flow passes per pair) -> 7x7-dilated
    # low-confidence mask -> cubic warp of the stylised key frame composited over the frame, 8 pairs per call
    from sd_animation_optical_flow_b200 import pdcnet_of
    from sd_animation_optical_flow_b200.engine import RaftFlowConfidence
    from tests import golden_inputs as gi
    algo3 = pdcnet_of.PDCNetPlus(network=RaftFlowConfidence(eng))
    canvas = gi.texture(H + 64, W + 64, 4242)
    clip = torch.from_numpy(np.stack([canvas[(3 * i) % 48:(3 * i) % 48 + H, (5 * i) % 56:(5 * i) % 56 + W] for i in range(32)])).to(dev)
    key8, sty8 = clip[:1].expand(P, -1, -1, -1).contiguous(), dsty

    def clip_pass():
        outs = []
        for i0 in range(1, 32, P):
            tgt = clip[i0:i0 + P]
            if tgt.shape[0] < P:                                    # keep one graph shape: pad the last call
                tgt = torch.cat([tgt, clip[-1:].expand(P - tgt.shape[0], -1, -1, -1)], 0)
            flow_c, wm_c = algo3._estimate(key8, tgt.contiguous())
            outs.append(ops.warp_mask_composite(sty8, tgt.contiguous(), flow_c, wm_c, 0.95, 7))
        return outs

    t_clip = time_op(clip_pass, 2, torch)
    clip_leg = {'workload': 'configs[2] shape: 32-frame 768x512 clip vs its key frame: flow + forward-backward confidence + dilated mask + composite',
                'pairs': 31, 'pairs_per_call': P, 'ms_per_clip': t_clip * 1e3, 'value': 31 / t_clip, 'unit': UNIT, 'n_gpus': 1,
                'note': 'two RAFT passes per pair (confidence by forward-backward consistency); device-resident'}

    cpu = cpu_baseline_leg(args.cpu_budget_s) if args.cpu_budget_s > 0 else None
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'tf32', 'data': 'synthetic',
            'config': {'workload': 'configs[1]: RAFT all-pairs correlation + warp, single 512x768 frame pair per GPU',
                       'H': H, 'W': W, 'iters': ITERS, 'weights': 'random-init (name-seeded)', 'corr_precision': args.corr_precision,
                       'conv_precision': 'bf16 autocast' if args.mixed_precision else 'cuDNN TF32 tensor cores, fp32 accumulate (torch default = what the reference runs on this GPU)',
                       'precision_note': 'dtype names the arithmetic of the bulk of the step (TF32 convolutions); the correlation volume uses '
                                         f'{args.corr_precision} operands with fp32 accumulation, the thin convolutions and all glue fp32, the warp exact integer (u8). '
                                         'Final flow vs the reference RAFT (fp32): EPE 3e-4 px mean with the fp16/tf32 volume, 1e-5 px with --corr-precision 3xtf32 '
                                         '(tests/test_gpu_raft.py)',
                       'cuda_graph': not args.no_graph, 'pairs_per_step_per_gpu': 1, 'parallelism': f'pairs x{world}, no collective',
                       'l2': 'per-step working set (200.5 MB pyramid rewritten every step + activations) exceeds the 126 MB L2; no explicit flush'},
            'roofline': roofline, 'roofline_extra': extra, 'batched': batched, 'clip': clip_leg, 'cpu_baseline': cpu, 'e2e': e2e, 'clocks': clocks,
            'gpu_launches': int(launches),
            'gpu_launches_note': f'{launches_per_step} libsdof_b200 kernels per step (counted on an eager step) x {args.steps} steps'
                                 + ('; CUDA-graph replays re-run the captured launches' if not args.no_graph else '')}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--corr-precision', default='fp16', choices=['fp16', 'tf32', '3xtf32', 'bf16', 'fp32'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--mixed-precision', action='store_true')
    ap.add_argument('--channels-last', action='store_true')
    ap.add_argument('--no-cudnn-benchmark', action='store_true')
    ap.add_argument('--no-side-streams', action='store_true')
    ap.add_argument('--cudnn-convf1', action='store_true')
    ap.add_argument('--cudnn-fh2', action='store_true')
    ap.add_argument('--quick', action='store_true', help='step timing only: skip e2e, roofline and CPU legs')
    ap.add_argument('--cpu-budget-s', type=float, default=10.0)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
