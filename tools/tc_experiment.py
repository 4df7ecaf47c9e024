import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import torch
    from sd_animation_optical_flow_b200 import ops
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev).manual_seed(0)
    f1 = torch.randn((1, 96, 64, 256), generator=g, device=dev)
    f2 = torch.randn((1, 96, 64, 256), generator=g, device=dev)
    for prec in ('fp16',):
        fn = lambda: ops.corr_volume_pyramid(f1, f2, 4, prec)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20): fn()
        e.record(); torch.cuda.synchronize()
        print(f'px={os.environ.get("SDOF_RES_PATCHX","-"):>2s} debug={os.environ.get("SDOF_RES_DEBUG","0"):>2s} {prec}: {s.elapsed_time(e)/20*1e3:.1f} us', flush=True)
else:
    for px in (32,):
        for dbg in (0, 13, 29, 16, 17, 4, 8):
            env = dict(os.environ, SDOF_RES_DEBUG=str(dbg), SDOF_RES_PATCHX=str(px))
            subprocess.run([sys.executable, __file__, 'child'], env=env, timeout=300)
