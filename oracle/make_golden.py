"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the authoring container).

    python -m oracle.make_golden            # needs /root/reference (read-only) and cv2

The reference is Python, so it is imported where it lies (RAFT/core via sys.path, exactly as
ofgen.py:57-58 does); the functions of ofgen_*.py / pdcnet_of.py that cannot be imported
(their modules import DenseMatching / Stable Diffusion at top level) are exercised through
the literal cv2/numpy expressions they consist of, quoted below with file:line.
Inputs are regenerated from seeds by tests/golden_inputs.py, only OUTPUTS are stored.
/root/reference does not exist on the GPU box, so nothing at test time imports it.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REF = os.environ.get('SDOF_REFERENCE', '/root/reference')

from tests import golden_inputs as gi  # noqa: E402


def ref_corr():
    import torch
    sys.path.insert(0, os.path.join(REF, 'RAFT', 'core'))
    from corr import CorrBlock  # reference RAFT/core/corr.py
    out = {}
    for name, (B, C, h, w) in gi.CORR_CASES.items():
        f1, f2, coords = gi.corr_inputs(name)
        cb = CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), num_levels=4, radius=4)
        look = cb(torch.from_numpy(coords)).numpy()
        out[f'{name}_lookup'] = look.astype(np.float32)
        for l, lv in enumerate(cb.corr_pyramid):
            rows = gi.PYRAMID_ROWS(lv.shape[0])
            out[f'{name}_pyr{l}'] = lv.numpy()[rows, 0].astype(np.float32)
    np.savez_compressed(os.path.join(GOLDEN, 'corr.npz'), **out)
    print('corr.npz', {k: v.shape for k, v in out.items()})


def ref_raft():
    import torch
    sys.path.insert(0, os.path.join(REF, 'RAFT', 'core'))
    from raft import RAFT  # reference RAFT/core/raft.py
    from utils.utils import InputPadder
    from sd_animation_optical_flow_b200.raft import fill_weights_by_name

    class namespace:  # ofgen.py:51-53
        def __contains__(self, m):
            return hasattr(self, m)

    out = {}
    for name, cfg in gi.RAFT_CASES.items():
        args = namespace()
        args.small = cfg['small']
        args.mixed_precision = False
        args.alternate_corr = False
        model = RAFT(args)
        fill_weights_by_name(model, cfg['seed'])
        model.eval()
        img1, img2 = gi.raft_inputs(name)  # RGB uint8 [H,W,3]
        t1 = torch.from_numpy(img1).permute(2, 0, 1).float()[None]
        t2 = torch.from_numpy(img2).permute(2, 0, 1).float()[None]
        padder = InputPadder(t1.shape)
        t1, t2 = padder.pad(t1, t2)
        with torch.no_grad():
            flow_low, flow_up = model(t1, t2, iters=cfg['iters'], test_mode=True)
        out[f'{name}_flow_low'] = flow_low[0].numpy().astype(np.float32)
        out[f'{name}_flow_up'] = flow_up[0].numpy().astype(np.float32)
        print(name, 'flow_up', flow_up.shape, 'mean |flow|', float(flow_up.abs().mean()))
    np.savez_compressed(os.path.join(GOLDEN, 'raft.npz'), **out)
    # state-dict keys/shapes of the reference models: the checkpoint-compatibility contract
    import json
    keys = {}
    for small in (False, True):
        args = namespace()
        args.small = small
        args.mixed_precision = False
        args.alternate_corr = False
        sd = RAFT(args).state_dict()
        keys['small' if small else 'basic'] = {k: list(v.shape) for k, v in sd.items()}
    with open(os.path.join(GOLDEN, 'raft_state_dict_keys.json'), 'w') as f:
        json.dump(keys, f, indent=0, sort_keys=True)


def ref_warp():
    import cv2
    out = {}
    for name in gi.WARP_CASES:
        img, flow = gi.warp_inputs(name)
        h, w = flow.shape[:2]
        # pdcnet_of.py:34-42
        X, Y = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
        map_x = (X + flow[:, :, 0]).astype(np.float32)
        map_y = (Y + flow[:, :, 1]).astype(np.float32)
        out[f'{name}_pdcnet'] = cv2.remap(img, map_x, map_y, interpolation=cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
        # ofgen.py:37-43
        f = -flow
        f[:, :, 0] += np.arange(w)
        f[:, :, 1] += np.arange(h)[:, np.newaxis]
        out[f'{name}_raft'] = cv2.remap(img, f, None, cv2.INTER_CUBIC)
    np.savez_compressed(os.path.join(GOLDEN, 'warp.npz'), **out)
    print('warp.npz', {k: (v.shape, v.dtype) for k, v in out.items()})


def ref_mask():
    import cv2
    out = {}
    conf, logc, img, raw, warped = gi.mask_inputs()
    # generate_mask, ofgen_pixel_inpaint.py:262-267
    for thres in (0.5, 0.95):
        mask = np.zeros((conf.shape[0], conf.shape[1]), dtype=np.uint8)
        mask[conf < thres] = 255
        lc = logc.copy()
        lc[conf < thres] = 0
        kern = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 7))
        out[f'mask_{thres}'] = cv2.dilate(mask, kern)
        out[f'logc_{thres}'] = lc
    m = out['mask_0.5']
    # confidence_to_mask's 15x15 dilation, ofgen_pixel_inpaint.py:219,226
    out['dilate15'] = cv2.dilate((conf < 0.2).astype(np.uint8) * 255, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (15, 15)))
    # ofgen_keyframe_inpaint.py:772-774
    out['invert_dilate'] = cv2.dilate(255 - m, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 7)))
    # expand_mask, ofgen_keyframe_inpaint.py:968-973
    lap = (cv2.cvtColor(np.absolute(cv2.Laplacian(img, cv2.CV_64F)).astype(np.uint8), cv2.COLOR_RGB2GRAY) > 20).astype(np.uint8) * 255
    lap = cv2.dilate(lap, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 7)))
    out['expand'] = cv2.bitwise_or(m, lap)
    # mix_propagated_ai_frame, ofgen_pixel_inpaint.py:251-260
    for ppw in (1.0, 0.3):
        weights = np.zeros((raw.shape[0], raw.shape[1]), dtype=np.float32)
        weights[m <= 127] = ppw
        weights[m > 127] = 1 - ppw
        weights = weights[:, :, None]
        ai = raw.astype(np.float32) * (1 - weights) + warped.astype(np.float32) * weights
        out[f'mix_{ppw}'] = np.clip(ai, 0, 255).astype(np.uint8)
    # merge_images naive, ofgen_keyframe_inpaint.py:676-681
    base = np.copy(raw)
    mask2 = (m / 255).astype(np.uint8)[:, :, None]
    out['merge'] = base * (1 - mask2) + warped * mask2
    # of_calc travel distance, ofgen_pixel_inpaint.py:105-116
    flow = gi.warp_inputs('small_u8')[1][: conf.shape[0], : conf.shape[1]].copy()
    h, w = flow.shape[:2]
    X, Y = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
    map_x = (X + flow[:, :, 0]).astype(np.float32)
    map_y = (Y + flow[:, :, 1]).astype(np.float32)
    map_x -= np.arange(w)
    map_y -= np.arange(h)[:, np.newaxis]
    v = np.sqrt(map_x * map_x + map_y * map_y)
    v[conf < 0.9] = 0
    out['travel'] = v
    np.savez_compressed(os.path.join(GOLDEN, 'mask.npz'), **out)
    print('mask.npz', {k: (v.shape, v.dtype) for k, v in out.items()})


def ref_greedy():
    """The greedy loop of ofgen_keyframe_inpaint.py:995-1024, quoted with cv2 as the warp."""
    import cv2
    flow_mat, frames, thres = gi.greedy_inputs()
    fm = flow_mat.copy()

    def warp_frame(frame, flow):  # pdcnet_of.py:34-42
        h, w = flow.shape[:2]
        X, Y = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
        return cv2.remap(frame, (X + flow[:, :, 0]).astype(np.float32), (Y + flow[:, :, 1]).astype(np.float32),
                         interpolation=cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)

    fm[:, :, :, :, 2] = (fm[:, :, :, :, 2] > thres).astype(np.float32)
    mask = np.zeros((fm.shape[2], fm.shape[3]), dtype=np.uint8)
    ret = None
    order = []
    for _ in range(fm.shape[0]):
        vals = fm[:, :, :, :, 2].reshape(fm.shape[0], -1).sum(axis=1)
        ref = int(np.argmax(vals))
        order.append(ref)
        warped = warp_frame(frames[ref], fm[ref, 0, :, :, 0:2])
        last = fm[ref, 0, :, :, 2]
        cur = (last * 255).astype(np.uint8)
        mask = cv2.bitwise_or(mask, cur)
        if ret is None:
            ret = np.copy(warped)
        else:
            m2 = (cur / 255).astype(np.uint8)[:, :, None]
            ret = ret * (1 - m2) + warped * m2
        fm[:, 0, :, :, 2] -= last[None, :, :]
        fm[:, 0, :, :, 2] = np.clip(fm[:, 0, :, :, 2], 0, 1)
    np.savez_compressed(os.path.join(GOLDEN, 'greedy.npz'), ret=ret, mask=mask, order=np.array(order), conf_after=fm[..., 2])
    print('greedy.npz order', order)


def ref_raft_full():
    """Full-size goldens of the configurations bench.py times: the reference RAFT (fp32, CPU) at
    768x512 (config 2/3/4 size, the bench pair of rank 0) and 720x1280 (config 5 size), iters=20 as in ofgen.py:77.
    flow_low is stored whole, flow_up on the lattice gi.full_lattice (one pixel per 8x8 block) to keep the fixture small."""
    import torch
    sys.path.insert(0, os.path.join(REF, 'RAFT', 'core'))
    from raft import RAFT  # reference RAFT/core/raft.py
    from sd_animation_optical_flow_b200.raft import fill_weights_by_name

    class namespace:  # ofgen.py:51-53
        def __contains__(self, m):
            return hasattr(self, m)

    out = {}
    for name, cfg in gi.RAFT_FULL_CASES.items():
        args = namespace()
        args.small = False
        args.mixed_precision = False
        args.alternate_corr = False
        model = RAFT(args)
        gi.raft_full_weights(model, name)
        model.eval()
        img1, img2 = gi.raft_full_inputs(name)
        t1 = torch.from_numpy(img1).permute(2, 0, 1).float()[None]
        t2 = torch.from_numpy(img2).permute(2, 0, 1).float()[None]
        with torch.no_grad():
            flow_low, flow_up = model(t1, t2, iters=cfg['iters'], test_mode=True)
        ys, xs = gi.full_lattice(*cfg['hw'])
        out[f'{name}_flow_low'] = flow_low[0].numpy().astype(np.float32)
        out[f'{name}_flow_up_s'] = np.ascontiguousarray(flow_up[0].numpy()[:, ys][:, :, xs]).astype(np.float32)
        print(name, 'flow_up', tuple(flow_up.shape), 'mean |flow|', float(flow_up.abs().mean()), 'max', float(flow_up.abs().max()))
    np.savez_compressed(os.path.join(GOLDEN, 'raft_full.npz'), **out)


if __name__ == '__main__':
    os.makedirs(GOLDEN, exist_ok=True)
    ref_warp()
    ref_mask()
    ref_greedy()
    ref_corr()
    ref_raft()
    ref_raft_full()
