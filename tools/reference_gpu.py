"""Reference-on-B200 baseline (SURVEY §8d "Reference-GPU baseline"): the UNMODIFIED reference RAFT
(baseline/_ref/RAFT/core, installed by baseline/install_ref.py) and its own compiled op (oracle/_ref/alt_cuda_corr_ref.so,
built by oracle/build_ref.py) timed with CUDA events on the same GPU, next to this package's kernels for the same inputs.

    python tools/reference_gpu.py            # prints one JSON object

`measure(dev)` is what bench.py calls for its `reference_gpu` key (rank 0, N=1).  Inputs as BASELINE.md §3.2:
fmaps ~ N(0,1) [1,256,96,64] and [1,256,90,160], coords = grid + 2*N(0,1), seed 0; RAFT with the name-seeded random
weights bench.py uses, iters=20, test_mode=True, on bench.py's synthetic pair.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _time(fn, n, torch, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e-3


def _import_reference():
    """The reference's RAFT/core modules, imported the way ofgen.py:57-58 does (sys.path -> RAFT/core), with its compiled
    op `alt_cuda_corr` (corr.py:5-9 imports it inside try/except) provided by the unmodified build in oracle/_ref."""
    from baseline import install_ref
    from oracle import build_ref
    core = install_ref.installed_core()
    if core is None:
        return None, None, 'baseline/_ref/RAFT/core not installed (python baseline/install_ref.py in the authoring container)'
    ref_op = None
    if build_ref.built_module_path() is not None:
        ref_op = build_ref.load()
        sys.modules['alt_cuda_corr'] = ref_op
    if core not in sys.path:
        sys.path.insert(0, core)
    import importlib
    mods = {name: importlib.import_module(name) for name in ('corr', 'raft')}
    return mods, ref_op, None


class _Namespace:  # ofgen.py:51-53
    def __contains__(self, m):
        return hasattr(self, m)


def measure(dev=None, sizes=((96, 64), (90, 160)), raft_sizes=((768, 512),), iters=20, quick=False):
    import numpy as np
    import torch

    from sd_animation_optical_flow_b200 import ops
    from sd_animation_optical_flow_b200.engine import RaftEngine
    from sd_animation_optical_flow_b200.raft import coords_grid, fill_weights_by_name
    from tests import golden_inputs as gi

    dev = torch.device('cuda', 0) if dev is None else dev
    mods, ref_op, why = _import_reference()
    if mods is None:
        return {'unavailable': why}
    CorrBlock, AlternateCorrBlock = mods['corr'].CorrBlock, mods['corr'].AlternateCorrBlock
    RefRAFT = mods['raft'].RAFT
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    out = {'note': 'unmodified reference code from baseline/_ref (RAFT/core) and oracle/_ref (alt_cuda_corr), CUDA events, same GPU',
           'corr': [], 'raft': []}
    try:
        with torch.no_grad():
            for (h, w) in sizes:
                g = torch.Generator(device=dev).manual_seed(0)
                f1 = torch.randn((1, 256, h, w), generator=g, device=dev)
                f2 = torch.randn((1, 256, h, w), generator=g, device=dev)
                coords = coords_grid(1, h, w, dev) + 2 * torch.randn((1, 2, h, w), generator=g, device=dev)
                row = {'fmap': [1, 256, h, w]}
                # --- reference CorrBlock: cuBLAS SGEMM (TF32 off: torch default for matmul, = what the reference runs) + scale + 3 avg_pool2d
                torch.backends.cuda.matmul.allow_tf32 = False
                holder = {}

                def build():
                    holder['cb'] = CorrBlock(f1, f2, num_levels=4, radius=4)
                row['ref_corrblock_build_us'] = _time(build, 10, torch) * 1e6
                cb = holder['cb']
                row['ref_corrblock_call_us'] = _time(lambda: cb(coords), 20, torch) * 1e6
                ref_look = cb(coords)
                del cb, holder
                # --- ours: same protocol (fp16 operands, the bench default) and the fp32-faithful mode
                n1 = h * w
                f1n, f2n = f1.permute(0, 2, 3, 1).contiguous(), f2.permute(0, 2, 3, 1).contiguous()
                for prec in ('fp16', '3xtf32'):
                    ph = {}

                    def ours_build(prec=prec):
                        ph['p'] = ops.corr_volume_pyramid(f1n, f2n, 4, prec)
                    row[f'ours_pyramid_{prec}_us'] = _time(ours_build, 10, torch) * 1e6
                    lo = torch.empty((1, 324, h, w), device=dev)
                    row[f'ours_lookup_{prec}_us'] = _time(lambda: ops.corr_lookup(ph['p'], coords, 4, out=lo), 20, torch) * 1e6
                    row[f'ours_lookup_{prec}_max_abs_err_vs_ref'] = float((ops.corr_lookup(ph['p'], coords, 4) - ref_look).abs().max())
                    del ph
                row['ratio_build_ref_over_ours_fp16'] = row['ref_corrblock_build_us'] / row['ours_pyramid_fp16_us']
                row['ratio_call_ref_over_ours_fp16'] = row['ref_corrblock_call_us'] / row['ours_lookup_fp16_us']
                # --- K1: the reference's compiled kernel vs sdof_alt_corr_forward, level 0 and the 4-level AlternateCorrBlock call
                if ref_op is not None:
                    c5 = coords.permute(0, 2, 3, 1).unsqueeze(1).contiguous()
                    row['ref_alt_cuda_corr_forward_us'] = _time(lambda: ref_op.forward(f1n, f2n, c5, 4), 10, torch) * 1e6
                    row['ours_alt_corr_forward_us'] = _time(lambda: ops.alt_corr_forward(f1n, f2n, c5, 4), 10, torch) * 1e6
                    a = ref_op.forward(f1n, f2n, c5, 4)[0]
                    b = ops.alt_corr_forward(f1n, f2n, c5, 4)
                    row['alt_corr_max_abs_diff'] = float((a - b).abs().max())
                    row['ratio_alt_ref_over_ours'] = row['ref_alt_cuda_corr_forward_us'] / row['ours_alt_corr_forward_us']
                    acb = AlternateCorrBlock(f1, f2, num_levels=4, radius=4)
                    row['ref_alternate_corrblock_call_us'] = _time(lambda: acb(coords), 5, torch) * 1e6
                    from sd_animation_optical_flow_b200 import corr as our_corr
                    oacb = our_corr.AlternateCorrBlock(f1, f2, num_levels=4, radius=4)
                    row['ours_alternate_corrblock_call_us'] = _time(lambda: oacb(coords), 5, torch) * 1e6
                    row['ratio_alternate_call_ref_over_ours'] = row['ref_alternate_corrblock_call_us'] / row['ours_alternate_corrblock_call_us']
                out['corr'].append({k: (round(v, 3) if isinstance(v, float) else v) for k, v in row.items()})
                del f1, f2, f1n, f2n
                torch.cuda.empty_cache()

            # --- the whole reference forward, as ofgen.py:62-78 runs it (fp32 module, torch default backends flags)
            for (H, W) in raft_sizes:
                f1, f2 = gi.shifted_pair(H, W, 1000)
                t1 = torch.from_numpy(f1).permute(2, 0, 1).float()[None].to(dev)
                t2 = torch.from_numpy(f2).permute(2, 0, 1).float()[None].to(dev)
                row = {'HxW': [H, W], 'iters': iters}
                for alt in ((False, True) if ref_op is not None else (False,)):
                    args = _Namespace()
                    args.small, args.mixed_precision, args.alternate_corr = False, False, alt
                    model = fill_weights_by_name(RefRAFT(args), 0, flow_head_scale=0.02).to(dev).eval()   # bench.py's weights
                    tag = 'alternate_corr' if alt else 'corrblock'
                    for conv_tf32 in ((True,) if quick else (True, False)):
                        torch.backends.cudnn.allow_tf32 = conv_tf32     # torch's default is True: what the unmodified reference runs
                        torch.backends.cuda.matmul.allow_tf32 = False   # torch's default
                        torch.backends.cudnn.benchmark = False          # the reference never sets it
                        torch.cuda.reset_peak_memory_stats(dev)
                        t = _time(lambda: model(t1, t2, iters=iters, test_mode=True), 3, torch, warm=2)
                        key = f'ref_raft_{tag}_{"tf32conv" if conv_tf32 else "fp32conv"}'
                        row[key + '_ms'] = round(t * 1e3, 3)
                        row[key + '_peak_mem_mb'] = round(torch.cuda.max_memory_allocated(dev) / 2 ** 20, 1)
                    if not alt:
                        flow_ref = model(t1, t2, iters=iters, test_mode=True)[1]
                    del model
                # ours on the same pair, same weights: device-resident estimate_flow (bench defaults)
                torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
                eng = RaftEngine(checkpoint=None, iters=iters, seed=0, flow_head_scale=0.02, device=dev)
                a = torch.from_numpy(f1).to(dev)[None]
                b = torch.from_numpy(f2).to(dev)[None]
                row['ours_estimate_flow_ms'] = round(_time(lambda: eng.estimate_flow(a, b), 10, torch, warm=3) * 1e3, 3)
                d = (eng.estimate_flow(a, b)[0].permute(2, 0, 1) - flow_ref[0]).norm(dim=0)
                row['ours_vs_ref_gpu_epe_mean_px'] = float(d.mean())
                row['ref_mean_abs_flow_px'] = float(flow_ref.abs().mean())
                row['ratio_ref_corrblock_tf32conv_over_ours'] = round(row['ref_raft_corrblock_tf32conv_ms'] / row['ours_estimate_flow_ms'], 2)
                out['raft'].append(row)
                del eng
                torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    return out


if __name__ == '__main__':
    print(json.dumps(measure(raft_sizes=((768, 512), (720, 1280)))))
