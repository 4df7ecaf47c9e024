"""Benchmark of the flow -> warp hot path (BASELINE.json metric: frame-pairs/sec (flow+warp) at 512x768;
corr-volume tensor-pipe % of peak).

    python bench.py --gpus 1 --steps 20 --warmup 3                      # our arm
    python bench.py --impl reference --gpus 1 --steps 5 --warmup 1      # CPU reference arm
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W       # N > 1, one rank per GPU

A "step" is one pass of the hot path over one synthetic 768x512 frame pair per rank (configs[1]: RAFT all-pairs
correlation + warp, iters=20): estimate_flow + cubic warp of the stylised previous frame.  Pairs are independent, so
ranks hold different pairs and there is no collective on the data path (weak scaling); `value` = pairs all ranks
processed / max-over-ranks device time.  The K steps are repeated back to back until the timed region is >= 1 s
(`timed_steps` of them, all timed, `ms_per_step` is their mean) so that the nvidia-smi clock samples fall inside it.

Prints ONE JSON line (rank 0).  Beside the contract keys: roofline (dominant hand-written kernel), roofline_extra,
cpu_baseline, e2e, clocks, gpu_launches, reference_gpu (the UNMODIFIED reference on this GPU, baseline/_ref), and the
other configs of BASELINE.json as legs: batched (8 pairs per call), clip (configs[2]), config4 (256-frame clip sharded
over the ranks), config5 (1000 x 720x1280 frames, key frame every 25: flow + warp + mask generation) and, for N > 1,
gather (the optional NCCL reassembly of the flow stack).
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

T_START = time.perf_counter()
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 768, 512          # tensors are H=768, W=512 (cv2.resize(frame,(512,768)), ofgen_pixel_inpaint.py:324)
ITERS = 20               # ofgen.py:77
METRIC = 'frame-pairs/sec (flow+warp) at 512x768'
UNIT = 'frame-pairs/s'
# Weights: name-seeded random init (no checkpoint ships with the reference) with the flow head's output convolution scaled
# by 0.02, so the flow stays a few px like a trained model's (plain random weights run away to ~200 px in 20 iterations and
# amplify TF32 rounding with it).  This is tests/golden_inputs.py::RAFT_FULL_CASES['S_calm'], for which the reference's own
# flow is committed (tests/golden/raft_full.npz): bench.py measures its EPE against it live (config.parity_epe).
WEIGHT_SEED, FLOW_HEAD_SCALE = 0, 0.02
MIN_TIMED_S = 1.0


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
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi sampling DURING the timed regions (B200_PROFILING.md recipe), 100 ms period; `mark()` notes the sample
    index at the start / end of a region so the medians are taken over samples INSIDE the regions."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.marks = []

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def _count(self):
        try:
            self.f.flush()
            return sum(1 for _ in open(self.path))
        except Exception:
            return 0

    def mark(self):
        if self.proc is not None:
            self.marks.append(self._count())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                rows.append((float(parts[1]), float(parts[2]), [n for n, v in zip(names, parts[5:9]) if v.lower().startswith('active')]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = []
        for a, b in zip(self.marks[0::2], self.marks[1::2]):
            inside += rows[a:b + 1]
        use = inside if inside else rows
        reasons = sorted({r for row in use for r in row[2]})
        return {'sm_mhz': statistics.median([r[0] for r in use]) if use else None, 'sm_max_mhz': max([r[1] for r in use]) if use else None,
                'samples': len(rows), 'samples_in_timed_regions': len(inside), 'reasons': reasons}


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


def time_graphed(fn, reps: int, torch, rounds: int = 3):
    """Device time of fn() replayed `reps` times inside ONE CUDA graph (no host launch cost between kernels): for kernels
    shorter than the host-side cost of launching them through Python."""
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
        best = min(best, s.elapsed_time(e) / reps * 1e-3)
    return best


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from sd_animation_optical_flow_b200 import _capi, ofgen, ops, shard
    from sd_animation_optical_flow_b200 import build as _build
    from sd_animation_optical_flow_b200.engine import RaftEngine
    from tests import golden_inputs as gi

    t_import = time.perf_counter()
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
    t_init = time.perf_counter()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    f1, f2, sty = synthetic_pair(rank)
    d1 = torch.from_numpy(f1).to(dev)[None]
    d2 = torch.from_numpy(f2).to(dev)[None]
    dsty = torch.from_numpy(sty).to(dev)[None]
    eng = RaftEngine(checkpoint=None, iters=ITERS, corr_precision=args.corr_precision, use_cuda_graph=not args.no_graph,
                     mixed_precision=args.mixed_precision, channels_last=args.channels_last, device=dev, seed=WEIGHT_SEED,
                     flow_head_scale=FLOW_HEAD_SCALE, cudnn_benchmark=not args.no_cudnn_benchmark,
                     fast_options=dict(side_streams=not args.no_side_streams, own_convf1=not args.cudnn_convf1, own_fh2=not args.cudnn_fh2,
                                       corr_storage=args.corr_storage, tc_gru=args.tc_gru, fnet_fp16=not args.no_fnet_fp16, cnet_fp16=not args.no_cnet_fp16, loop_fp16=not args.no_loop_fp16, defer_coords=not args.no_defer_coords, convf1_gemm=not args.no_convf1_gemm))

    def step():
        flow = eng.estimate_flow(d1, d2)                 # [1,768,512,2]
        return ops.warp(dsty, flow, 'cv2_cubic', -1.0)   # ofgen.warp_frame: previous stylised frame at x - flow

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_w0 = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    t_warm = time.perf_counter()
    # kernels of THIS library per step, counted on one eager (non-graph) step: graph replays re-run
    # exactly these launches without going through the host-side counter
    graphed = eng.use_cuda_graph
    eng.use_cuda_graph = False
    c0 = _capi.launch_count()
    step()
    launches_per_step = _capi.launch_count() - c0
    eng.use_cuda_graph = graphed
    # which warp path does the headline step take? (tile counters of the tiled kernel)
    ops.warp_tile_stats(reset=True)
    step()
    warp_staged, warp_fallback = ops.warp_tile_stats(reset=True)
    # K steps, repeated until the timed region is >= MIN_TIMED_S (every repetition is timed)
    est = time_op(step, 5, torch)
    reps = max(1, int(np.ceil(MIN_TIMED_S / max(est * args.steps, 1e-6))))
    if world > 1:
        t = torch.tensor([reps], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        reps = int(t.item())
    timed_steps = args.steps * reps
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    s.record()
    for _ in range(timed_steps):
        out = step()
    e.record()
    barrier()
    sampler.mark()
    dt = max_over_ranks(s.elapsed_time(e) * 1e-3)
    launches = launches_per_step * timed_steps
    value = world * timed_steps / dt
    if args.quick:
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            print(json.dumps({'quick': True, 'value': value, 'ms_per_step': dt / timed_steps * 1e3, 'timed_steps': timed_steps,
                              'launches_per_step': launches_per_step, 'clocks': clocks, 'argv': sys.argv[1:]}))
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
    sampler.mark()
    t0 = time.perf_counter()
    for _ in range(timed_steps):
        warped_np = e2e_step()
    torch.cuda.synchronize()
    e2e_dt = max_over_ranks(time.perf_counter() - t0)
    sampler.mark()
    clocks = sampler.stop() if rank == 0 else None   # sampled inside both timed regions (device-resident steps, e2e steps)
    frame_b, flow_b = H * W * 3, H * W * 2 * 4
    e2e = {'value': world * timed_steps / e2e_dt, 'unit': UNIT, 'h2d_bytes_per_step': 3 * frame_b + flow_b,
           'd2h_bytes_per_step': flow_b + frame_b, 'api': 'ofgen.RAFT_2.calc(np,np) + ofgen.warp_frame(np,np)', 'steps': timed_steps}

    # ---- configs[3]: a 256-frame 768x512 clip, consecutive-frame pairs (ofgen.py: flow prev -> cur, warp the previous
    # stylised frame) sharded contiguously over the ranks with shard.shard_range, 8 pairs per call, no collective
    P = 8
    canvas = torch.from_numpy(gi.texture(H + 96, W + 96, 4242)).to(dev)

    def clip_frame(i):     # smooth random-walk crop of one canvas: frame i of the synthetic clip
        oy = int(48 + 40 * np.sin(0.071 * i) + 6 * np.sin(0.9 * i))
        ox = int(48 + 40 * np.cos(0.053 * i) + 6 * np.cos(1.1 * i))
        return canvas[oy:oy + H, ox:ox + W]

    n_frames4 = 256
    p0, p1 = shard.shard_range(n_frames4 - 1, rank, world)   # pair j = (frame j, frame j+1)
    sty8 = dsty.expand(P, -1, -1, -1).contiguous()

    # flow stack of this rank's pairs (only kept when it is gathered afterwards); allocated BEFORE the timed pass
    local_stack = torch.empty((p1 - p0, H, W, 2), device=dev) if world > 1 else None

    def config4_pass(keep: bool = False):
        for j0 in range(p0, p1, P):
            js = list(range(j0, min(j0 + P, p1)))
            n_real = len(js)
            # keep one graph shape: the last call repeats its last frame (those pairs are dropped)
            # P + 1 consecutive frames -> P pairs: the feature encoder runs once per frame (estimate_flow_sequence)
            fr = torch.stack([clip_frame(j) for j in range(j0, j0 + n_real + 1)] + [clip_frame(j0 + n_real)] * (P - n_real))
            fl = eng.estimate_flow_sequence(fr)
            ops.warp(sty8, fl, 'cv2_cubic', -1.0)
            if keep and local_stack is not None:
                local_stack[j0 - p0:j0 - p0 + n_real].copy_(fl[:n_real])

    config4_pass()                                          # warm-up: graph capture of the 8-pair shape
    barrier()
    s4, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s4.record()
    config4_pass(keep=True)
    e4.record()
    barrier()
    t4 = max_over_ranks(s4.elapsed_time(e4) * 1e-3)
    config4 = {'workload': 'configs[3]: 256-frame 768x512 clip, 255 consecutive-frame pairs (flow via estimate_flow_sequence: fnet once per frame; + cubic warp), sharded contiguously over the ranks (shard.shard_range), no collective',
               'pairs': n_frames4 - 1, 'pairs_per_call': P, 'seconds': t4, 'value': (n_frames4 - 1) / t4, 'unit': UNIT, 'n_gpus': world}
    # the optional reassembly of the flow stack (north star: "NCCL over NVLink only for an optional gather")
    gather = None
    if world > 1:
        local = local_stack                                 # [pairs of this rank, H, W, 2] fp32
        shard.gather_stack(local[:1], world)                # NCCL warm-up (communicator, channels)
        del_me = shard.gather_stack(local, n_frames4 - 1)   # and one full-size call: the 2 x 0.8 GB of buffers come from cudaMalloc the first time
        del del_me
        barrier()
        sg, eg = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sg.record()
        full = shard.gather_stack(local, n_frames4 - 1)
        eg.record()
        barrier()
        tg = max_over_ranks(sg.elapsed_time(eg) * 1e-3)
        gather = {'what': 'shard.gather_stack: all_gather_into_tensor of the per-rank [pairs, 768, 512, 2] fp32 flow stacks (NCCL)',
                  'shape': list(full.shape), 'bytes_total': int(full.numel() * 4), 'seconds': tg,
                  'algbw_GBps': full.numel() * 4 / tg / 1e9}
        del full, local
    del local_stack
    torch.cuda.empty_cache()

    # ---- configs[4]: 1000 frames 720x1280, key frame every 25 (40 keys x 24 frames = 960 pairs): per key ONE encode_key
    # (fnet(key) + pooled correlation operands), then per 8 frames: keyed flow (image1 = frame, image2 = key), cubic warp of
    # the stylised key, confidence -> generate_mask (M3) -> expand_mask (M6) -> composite (M4).  Keys are sharded
    # contiguously over the ranks (every pair of a key stays on one rank), no collective.  RAFT has no confidence head:
    # seeded synthetic logits stand in for PDCNet+'s weight_map (SURVEY §8a note).
    H5, W5, KEY_EVERY, N5 = 720, 1280, 25, 1000
    canvas5 = torch.from_numpy(gi.texture(H5 + 96, W5 + 96, 777)).to(dev)

    def frame5(i):
        oy = int(48 + 40 * np.sin(0.031 * i) + 5 * np.sin(0.7 * i))
        ox = int(48 + 40 * np.cos(0.027 * i) + 5 * np.cos(0.8 * i))
        return canvas5[oy:oy + H5, ox:ox + W5]

    n_keys = N5 // KEY_EVERY
    k0, k1 = shard.shard_range(n_keys, rank, world)
    g5 = torch.Generator(device=dev).manual_seed(5)
    wm5 = torch.randn((P, 2, H5, W5), generator=g5, device=dev) * 2
    sty5 = torch.from_numpy(gi.texture(H5, W5, 778)).to(dev)[None]

    def config5_pass(keys):
        for k in keys:
            kf = eng.encode_key(frame5(k * KEY_EVERY).contiguous())
            for f0 in range(1, KEY_EVERY, P):
                fr = torch.stack([frame5(k * KEY_EVERY + f) for f in range(f0, f0 + P)])
                fl = eng.estimate_flow_keyed(kf, fr)
                warped = ops.warp(sty5, fl, 'cv2_cubic', 1.0)
                conf, logc = ops.confidence_softmax(wm5)
                mask = ops.generate_mask(conf, logc, 0.95, 7)
                mask = ops.expand_mask(mask, fr, 7)
                ops.mix_propagated(fr, warped, mask, 1.0)

    config5_pass([k0] if k1 > k0 else [])                  # warm-up: graph capture of the keyed 8-frame shape
    barrier()
    s5, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s5.record()
    config5_pass(range(k0, k1))
    e5.record()
    barrier()
    t5 = max_over_ranks(s5.elapsed_time(e5) * 1e-3)
    pairs5 = n_keys * (KEY_EVERY - 1)
    config5 = {'workload': 'configs[4]: 1000 frames 720x1280, key frame every 25: per key encode_key once, then keyed flow + cubic warp + generate_mask + expand_mask + composite, 8 frames per call; keys sharded contiguously over the ranks',
               'pairs': pairs5, 'keys': n_keys, 'pairs_per_call': P, 'seconds': t5, 'value': pairs5 / t5, 'unit': UNIT, 'n_gpus': world,
               'note': 'confidence = softmax of seeded synthetic logits (RAFT has no confidence head; PDCNet+ is absent)'}
    del canvas5, wm5
    torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity of THIS configuration (weights, pair, numerics) against the reference's flow for it
    parity = None
    gpath = os.path.join(ROOT, 'tests', 'golden', 'raft_full.npz')
    if os.path.exists(gpath) and args.corr_precision == 'fp16' and FLOW_HEAD_SCALE == gi.RAFT_FULL_CASES['S_calm']['fh_scale']:
        gold = np.load(gpath)
        i1, i2 = gi.raft_full_inputs('S_calm')              # == synthetic_pair(0)
        fl = eng.estimate_flow(torch.from_numpy(i1).to(dev)[None], torch.from_numpy(i2).to(dev)[None], unpad=False)[0]
        ys, xs = gi.full_lattice(H, W)
        got = fl.permute(2, 0, 1).cpu().numpy()[:, ys][:, :, xs]
        epe = np.sqrt(((got - gold['S_calm_flow_up_s']) ** 2).sum(0))
        parity = {'epe_mean_px': float(epe.mean()), 'epe_max_px': float(epe.max()),
                  'mean_flow_px': float(np.sqrt((gold['S_calm_flow_up_s'] ** 2).sum(0)).mean()),
                  'against': 'tests/golden/raft_full.npz S_calm: the reference RAFT/core/raft.py (fp32, CPU) on the same weights and pair, flow_up on one pixel per 8x8 block',
                  'tolerance_px': 1e-2}

    # ---- roofline of the hand-written kernels, timed alone with CUDA events on the launching stream
    peaks = load_peaks()
    n1 = (H // 8) * (W // 8)
    C = 256
    g = torch.Generator(device=dev).manual_seed(0)
    fm1 = torch.randn((1, H // 8, W // 8, C), generator=g, device=dev)
    fm2 = torch.randn((1, H // 8, W // 8, C), generator=g, device=dev)
    storage = eng.fast.corr_storage if eng.fast is not None else 'fp32'
    ebytes = 2 if storage == 'fp16' else 4
    if args.corr_precision in ('fp16', 'bf16'):
        src_ops, tgt_ops = ops.CorrSource(fm1, args.corr_precision), ops.CorrTarget(fm2, 4, args.corr_precision)
        pyr = src_ops.pyramid(tgt_ops, storage)
        # the dominant kernel alone (operands prepared): replayed inside a CUDA graph, as it runs in the step
        t_corr = time_graphed(lambda: src_ops.pyramid(tgt_ops, storage, out=pyr), 20, torch)
        t_prep = time_graphed(lambda: ops.prepare_pair(fm1, fm2, 4, args.corr_precision), 20, torch)
        corr_kernel = f'corr_pyramid_resident_kernel<{"fp16" if storage == "fp16" else "fp32"}-stored pyramid>'
        lay = pyr.layout
        pooled = sum(lay.h[l] * lay.w[l] for l in range(4))
        in_bytes = (n1 + pooled) * C * 2            # 16-bit operands: fmap1 + the pooled fmap2 levels
    else:
        pyr = ops.corr_volume_pyramid(fm1, fm2, 4, args.corr_precision)
        t_corr = time_graphed(lambda: ops.corr_volume_pyramid(fm1, fm2, 4, args.corr_precision), 10, torch)
        t_prep = 0.0
        corr_kernel = 'corr_volume_tc_kernel + operand pre-pass'
        lay = pyr.layout
        in_bytes = 2 * n1 * C * 4
    out_bytes = ebytes * n1 * sum(lay.h[l] * lay.w[l] for l in range(4))
    corr_flops = 2.0 * n1 * n1 * C
    from sd_animation_optical_flow_b200.raft import coords_grid
    coords = coords_grid(1, H // 8, W // 8, dev) + 2 * torch.randn((1, 2, H // 8, W // 8), generator=g, device=dev)
    look_out = torch.empty((1, 324, H // 8, W // 8), device=dev)
    t_look = time_graphed(lambda: ops.corr_lookup(pyr, coords, 4, out=look_out), 20, torch)
    coords_nhwc = coords.permute(0, 2, 3, 1).contiguous()
    look_nhwc = torch.empty((1, H // 8, W // 8, 324), device=dev)
    t_look_nhwc = time_graphed(lambda: ops.corr_lookup_nhwc(pyr, coords_nhwc, 4, look_nhwc), 20, torch)
    look_bytes = n1 * (4 * 100 * ebytes + 324 * 4)
    # flows as the path produces them: smooth fields (camera / object motion): per frame a random translation of a few
    # pixels plus a low-frequency deformation (1/64-resolution Gaussian field of 4 px, bicubic-upsampled: |grad| ~ 0.1)
    flow32 = (torch.nn.functional.interpolate(torch.randn((32, 2, H // 64, W // 64), generator=g, device=dev) * 4, scale_factor=64,
                                              mode='bicubic', align_corners=False)
              + 6 * torch.randn((32, 2, 1, 1), generator=g, device=dev)).permute(0, 2, 3, 1).contiguous()
    src32 = torch.randint(0, 256, (32, H, W, 3), dtype=torch.uint8, device=dev)
    t_warp = time_op(lambda: ops.warp(src32, flow32), 20, torch)
    ops.warp_tile_stats(reset=True)
    ops.warp(src32, flow32)
    w32_staged, w32_fallback = ops.warp_tile_stats(reset=True)
    wm32 = torch.randn((32, 2, H, W), generator=g, device=dev) * 3
    t_fused = time_op(lambda: ops.warp_mask_composite(src32[:1], src32, flow32, wm32, 0.95, 7), 20, torch)
    # TF32 tensor peak, measured like MEASURED_PEAKS.json's bf16 figure (torch.matmul 8192^3, best of 5): the step's
    # convolutions run in TF32, so the whole-step tensor fraction is quoted against THIS number
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    ma = torch.randn((8192, 8192), device=dev)
    mb = torch.randn((8192, 8192), device=dev)
    tf32_peak = max(2 * 8192.0 ** 3 / time_op(lambda: torch.matmul(ma, mb), 3, torch) / 1e12 for _ in range(5))
    torch.backends.cuda.matmul.allow_tf32 = old_tf32
    del ma, mb
    step_flops = 946e9           # reference RAFT at 768x512, iters=20 (torch.utils.flop_counter)
    fp16_step = not (args.no_loop_fp16 or args.tc_gru or args.mixed_precision)

    hbm = peaks['hbm_gbs']
    corr_gbs = (in_bytes + out_bytes) / t_corr / 1e9
    roofline = {'kernel': corr_kernel + ', 1 pair, N=6144, C=256, 4 levels', 'bound': 'hbm',
                'achieved': corr_gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': corr_gbs / hbm,
                'traffic': None,
                'traffic_note': 'not measured in this run (VERDICT r1: no constants here); one ncu --set full capture of the same kernel at this size is committed as profiles/r2_kernels_ncu_summary.txt: dram__bytes_read.sum 8.03 MB + dram__bytes_write.sum 48.94 MB -- below the algorithmic bytes because about half of the 100 MB pyramid is still dirty in the 126 MB L2 when the kernel ends; inside a step (profiles/r2_lookup_dram_in_step.txt) 0.45 MB read + 101.9 MB written',
                'peak_source': peaks['source'], 'us_per_launch': t_corr * 1e6, 'us_operand_prepass': t_prep * 1e6,
                'frac_with_prepass': (in_bytes + out_bytes + 2 * n1 * C * 4) / (t_corr + t_prep) / 1e9 / hbm,
                'algorithmic_bytes': in_bytes + out_bytes,
                'algorithmic_bytes_note': f'16-bit operands in ({in_bytes} B) + {ebytes}-byte pyramid out ({out_bytes} B); SURVEY §8(d) counts an fp32 pyramid (200.5 MB): with it the same launch would read {(2 * n1 * C * 4 + 4 * n1 * pooled if args.corr_precision in ("fp16", "bf16") else in_bytes + out_bytes) / t_corr / 1e9 / hbm:.2f} of peak, which is not what this kernel moves',
                'timing': 'CUDA events around 20 replays inside one CUDA graph (the kernel is shorter than its Python launch path)'}
    tf = corr_flops / t_corr / 1e12
    extra = [
        {'kernel': corr_kernel, 'bound': 'tensor', 'achieved': tf, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
         'frac': tf / peaks['bf16_tflops'], 'note': f'{args.corr_precision} MMA (2*N^2*C flops of level 0) vs measured bf16 burst peak'},
        {'kernel': f'corr_lookup_kernel<4,4,32> (planar output, the corr_fn protocol; {storage} pyramid)', 'bound': 'hbm', 'achieved': look_bytes / t_look / 1e9,
         'peak': hbm, 'unit': 'GB/s', 'frac': look_bytes / t_look / 1e9 / hbm, 'us_per_launch': t_look * 1e6, 'algorithmic_bytes': look_bytes},
        {'kernel': f'corr_lookup_kernel<4,4,8> (channels-last output, the one in the step; {storage} pyramid)', 'bound': 'hbm',
         'achieved': look_bytes / t_look_nhwc / 1e9, 'peak': hbm, 'unit': 'GB/s', 'frac': look_bytes / t_look_nhwc / 1e9 / hbm,
         'us_per_launch': t_look_nhwc * 1e6, 'algorithmic_bytes': look_bytes},
        {'kernel': 'warp_cubic_u8c3_tiled_kernel, 32 frames, smooth flow (translation + low-frequency deformation)', 'bound': 'hbm', 'achieved': 14.0 * 32 * H * W / t_warp / 1e9, 'peak': hbm,
         'unit': 'GB/s', 'frac': 14.0 * 32 * H * W / t_warp / 1e9 / hbm, 'us_per_launch': t_warp * 1e6,
         'tiles_staged': w32_staged, 'tiles_fallback': w32_fallback},
        {'kernel': 'warp_mask_composite_kernel, 32 frames, same flows, N(0,3) logits, thres 0.95, 7x7 ellipse', 'bound': 'hbm', 'achieved': 26.0 * 32 * H * W / t_fused / 1e9, 'peak': hbm,
         'unit': 'GB/s', 'frac': 26.0 * 32 * H * W / t_fused / 1e9 / hbm, 'us_per_launch': t_fused * 1e6},
        {'kernel': 'whole step (cuDNN tensor-op convolutions dominate: fp16 operands in the encoders and the update block)', 'bound': 'tensor',
         'achieved': step_flops / (dt / timed_steps) / 1e12, 'peak': peaks['bf16_tflops'] if fp16_step else tf32_peak,
         'unit': 'TFLOP/s', 'frac': step_flops / (dt / timed_steps) / 1e12 / (peaks['bf16_tflops'] if fp16_step else tf32_peak),
         'frac_of_tf32_peak': step_flops / (dt / timed_steps) / 1e12 / tf32_peak,
         'note': '946 GFLOP per pair (reference RAFT, 768x512, iters=20) / step time vs the measured 16-bit dense peak (MEASURED_PEAKS.json) when the convolutions run on fp16 operands, else vs the TF32 matmul peak measured in this run (torch.matmul 8192^3, allow_tf32); at 6144 pixels per pair every convolution of the update loop is a sub-wave GEMM bound by launch latency and L2 -> SM operand traffic, not by the tensor pipe'},
    ]

    # ---- the same path with 8 pairs per call (how configs[2..4] feed it: PDCNetAux batches 16 pairs,
    # ofgen_keyframe_inpaint.py:550,585-600); reported beside the single-pair headline, not instead of it
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
    algo3 = pdcnet_of.PDCNetPlus(network=RaftFlowConfidence(eng))
    clip = torch.stack([clip_frame(3 * i) for i in range(32)])
    key8, sty1 = clip[:1].expand(P, -1, -1, -1).contiguous(), dsty

    def clip_pass():
        outs = []
        for i0 in range(1, 32, P):
            tgt = clip[i0:i0 + P]
            if tgt.shape[0] < P:                                    # keep one graph shape: pad the last call
                tgt = torch.cat([tgt, clip[-1:].expand(P - tgt.shape[0], -1, -1, -1)], 0)
            flow_c, wm_c = algo3._estimate(key8, tgt.contiguous())
            outs.append(ops.warp_mask_composite(sty1, tgt.contiguous(), flow_c, wm_c, 0.95, 7))
        return outs

    t_clip = time_op(clip_pass, 2, torch)
    clip_leg = {'workload': 'configs[2] shape: 32-frame 768x512 clip vs its key frame: flow + forward-backward confidence + dilated mask + composite',
                'pairs': 31, 'pairs_per_call': P, 'ms_per_clip': t_clip * 1e3, 'value': 31 / t_clip, 'unit': UNIT, 'n_gpus': 1,
                'note': 'two RAFT passes per pair (confidence by forward-backward consistency); device-resident'}

    # ---- the UNMODIFIED reference on this GPU (baseline/_ref + oracle/_ref), N = 1 only
    reference_gpu = None
    if world == 1 and not args.no_reference_gpu:
        try:
            sys.path.insert(0, os.path.join(ROOT, 'tools'))
            import reference_gpu as rg
            reference_gpu = rg.measure(dev, quick=True)
        except Exception as ex:  # noqa: BLE001  (the reference baseline must never take the bench line down)
            reference_gpu = {'unavailable': f'{type(ex).__name__}: {ex}'}

    cpu = cpu_baseline_leg(args.cpu_budget_s) if args.cpu_budget_s > 0 else None
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': dt / timed_steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'tf32' if (args.no_loop_fp16 or args.tc_gru) else 'fp16', 'data': 'synthetic',
            'timed_steps': timed_steps, 'timed_region_s': dt,
            'timed_note': f'the {args.steps} steps are repeated {reps}x back to back so that the timed region is >= {MIN_TIMED_S} s; every repetition is inside the timed region',
            'config': {'workload': 'configs[1]: RAFT all-pairs correlation + warp, single 512x768 frame pair per GPU',
                       'H': H, 'W': W, 'iters': ITERS,
                       'weights': f'random-init, name-seeded (seed {WEIGHT_SEED}), flow-head output convolution x{FLOW_HEAD_SCALE} so the flow stays a few px (RAFT_FULL_CASES S_calm)',
                       'corr_precision': args.corr_precision, 'corr_storage': storage,
                       'conv_precision': 'bf16 autocast' if args.mixed_precision else ('cuDNN tensor cores, fp32 accumulate: TF32 (torch default = what the reference runs on this GPU) for what is left in fp32 activations (per-pair context maps)' + ('; fp16 activations / filters (the same 11-bit operand precision, fp32 accumulation) for the encoders and the update block, fp32 hidden-state master copy / coordinates / flow' if not args.no_loop_fp16 else '; fp16 encoders')),
                       'precision_note': 'dtype names the operand type of the bulk of the step: fp16 activations and filters with fp32 accumulation (11-bit significand, what TF32 - the arithmetic the reference itself gets on this GPU from torch defaults - rounds its operands to), fp32 recurrent state; config.parity_epe is the measured distance to the fp32 reference.  The correlation volume uses '
                                         f'auto-ranged {args.corr_precision} operands (11-bit significand like TF32, per-tensor power-of-two scale) with fp32 accumulation and a {storage}-stored pyramid, '
                                         'the thin convolutions and all glue fp32, the warp exact integer (u8)',
                       'parity_epe': parity,
                       'warp_path_of_the_step': {'tiles_staged': warp_staged, 'tiles_fallback': warp_fallback},
                       'cuda_graph': not args.no_graph, 'pairs_per_step_per_gpu': 1, 'parallelism': f'pairs x{world}, no collective',
                       'l2': 'per-step working set (pyramid rewritten every step + activations of 20 iterations) exceeds the 126 MB L2; no explicit flush'},
            'roofline': roofline, 'roofline_extra': extra, 'tf32_matmul_peak_tflops': tf32_peak,
            'batched': batched, 'clip': clip_leg, 'config4': config4, 'config5': config5, 'gather': gather,
            'reference_gpu': reference_gpu, 'cpu_baseline': cpu, 'e2e': e2e, 'clocks': clocks,
            'gpu_launches': int(launches),
            'gpu_launches_note': f'{launches_per_step} libsdof_b200 kernels per step (counted on an eager step) x {timed_steps} timed steps'
                                 + ('; CUDA-graph replays re-run the captured launches' if not args.no_graph else ''),
            'setup_s': {'python_imports': round(t_import - T_START, 2), 'process_group_build_load': round(t_init - t_import, 2),
                        'engine_construction': round(t_w0 - t_init, 2), 'warmup_incl_cudnn_autotune_and_graph_capture': round(t_warm - t_w0, 2),
                        'total_wall': round(time.perf_counter() - T_START, 2)}}
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
    ap.add_argument('--corr-storage', default=None, choices=['fp16', 'fp32'], help='pyramid storage (default: fp16 with 16-bit operands)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--mixed-precision', action='store_true')
    ap.add_argument('--channels-last', action='store_true')
    ap.add_argument('--no-cudnn-benchmark', action='store_true')
    ap.add_argument('--no-side-streams', action='store_true')
    ap.add_argument('--cudnn-convf1', action='store_true')
    ap.add_argument('--cudnn-fh2', action='store_true')
    ap.add_argument('--tc-gru', action='store_true', help='SepConvGRU on the tcgen05 kernels of csrc/conv_tc.cu instead of cuDNN + glue (A/B switch)')
    ap.add_argument('--no-fnet-fp16', action='store_true', help='feature encoder in TF32 (fp32 activations) instead of fp16 (A/B switch)')
    ap.add_argument('--no-cnet-fp16', action='store_true', help='context encoder in TF32 instead of fp16 (A/B switch)')
    ap.add_argument('--no-convf1-gemm', action='store_true', help='convf1 as the hand-written fp32 FMA kernel instead of im2col + cuDNN 1x1 tensor-core convolution (A/B switch)')
    ap.add_argument('--no-defer-coords', action='store_true', help='separate coords-update kernel after the flow head instead of applying it in the next lookup / convf1 (A/B switch)')
    ap.add_argument('--no-loop-fp16', action='store_true', help='update block in TF32 with fp32 activations instead of fp16 (A/B switch)')
    ap.add_argument('--no-reference-gpu', action='store_true', help='skip the reference-on-this-GPU leg')
    ap.add_argument('--quick', action='store_true', help='step timing only: skip e2e, roofline, config and CPU legs')
    ap.add_argument('--cpu-budget-s', type=float, default=10.0)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
