"""EPE of RaftEngine against the reference RAFT's full-size goldens (tests/golden/raft_full.npz) under several numeric
configurations: which of them meets the 1e-2 px bar of SURVEY §8(d) at the sizes bench.py times.

    python tools/parity_probe.py            # one JSON line per (case, configuration)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import golden_inputs as gi  # noqa: E402


def engine_for(name, dev, tmpdir, **kw):
    from sd_animation_optical_flow_b200.engine import RaftEngine
    from sd_animation_optical_flow_b200.raft import RAFT
    from types import SimpleNamespace
    cfg = gi.RAFT_FULL_CASES[name]
    m = RAFT(SimpleNamespace(small=False, mixed_precision=False, alternate_corr=False, corr_precision='fp16'))
    gi.raft_full_weights(m, name)
    path = os.path.join(tmpdir, f'{name}.pth')
    torch.save({'module.' + k: v for k, v in m.state_dict().items()}, path)   # the public checkpoints' key format
    return RaftEngine(checkpoint=path, iters=cfg['iters'], device=dev, **kw)


def epe_vs_golden(eng, name, golden, bgr=True):
    img1, img2 = gi.raft_full_inputs(name)
    dev = eng.device
    a = torch.from_numpy(np.ascontiguousarray(img1[:, :, ::-1] if bgr else img1)).to(dev)[None]
    b = torch.from_numpy(np.ascontiguousarray(img2[:, :, ::-1] if bgr else img2)).to(dev)[None]
    flow = eng.estimate_flow(a, b, unpad=False, bgr=bgr)[0].permute(2, 0, 1).cpu().numpy()
    H, W = gi.RAFT_FULL_CASES[name]['hw']
    ys, xs = gi.full_lattice(H, W)
    ref = golden[f'{name}_flow_up_s']
    got = flow[:, ys][:, :, xs]
    epe = np.sqrt(((got - ref) ** 2).sum(0))
    mag = np.sqrt((ref ** 2).sum(0))
    return {'epe_mean': float(epe.mean()), 'epe_max': float(epe.max()), 'flow_mean': float(mag.mean()),
            'rel_mean': float(epe.mean() / max(mag.mean(), 1e-9))}


def main():
    import tempfile
    dev = torch.device('cuda', 0)
    golden = np.load(os.path.join(ROOT, 'tests', 'golden', 'raft_full.npz'))
    tmp = tempfile.mkdtemp()
    names = sys.argv[1:] or list(gi.RAFT_FULL_CASES)
    for name in names:
        for tag, tf32, kw in (('bench defaults: tf32 convs, fp16 volume, graph', True, {}),
                              ('fp32 convs, fp16 volume', False, {}),
                              ('tf32 convs, 3xtf32 volume', True, dict(corr_precision='3xtf32')),
                              ('fp32 convs, 3xtf32 volume', False, dict(corr_precision='3xtf32')),
                              ('fp32 convs, 3xtf32 volume, module forward', False, dict(corr_precision='3xtf32', fast=False))):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            eng = engine_for(name, dev, tmp, **kw)
            r = epe_vs_golden(eng, name, golden)
            print(json.dumps({'case': name, 'config': tag, **{k: round(v, 6) for k, v in r.items()}}), flush=True)
            del eng
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
