"""Drop-in for the flow / warp / mask / composite helpers of the reference's
`ofgen.py`, `ofgen_pixel_inpaint.py` and `ofgen_keyframe_inpaint.py`: same names, numpy in /
numpy out, arithmetic on the B200.  (The Stable-Diffusion orchestration around them is out of
scope)

| here                      | reference                                                      |
|---------------------------|----------------------------------------------------------------|
| warp_frame                | ofgen.py:37-43 (x - flow; dups ofgen_pixel_inpaint.py:84-90)    |
| warp_frame_pdcnet         | pdcnet_of.py:34-42 as imported at ofgen_pixel_inpaint.py:17     |
| RAFT_2, create_of_algo    | ofgen.py:55-83                                                  |
| of_calc                   | ofgen.py:45-49 (RAFT flavour)                                   |
| of_calc_pdcnet            | ofgen_pixel_inpaint.py:105-118                                  |
| generate_mask             | ofgen_pixel_inpaint.py:262-267, ofgen_keyframe_inpaint.py:317-322 |
| mix_propagated_ai_frame   | ofgen_pixel_inpaint.py:251-260                                  |
| confidence_to_mask        | ofgen_pixel_inpaint.py:218-227                                  |
| merge_images              | ofgen_keyframe_inpaint.py:676-681 (method='naive')              |
| expand_mask               | ofgen_keyframe_inpaint.py:968-973                               |
| composite_references      | the greedy loop of ofgen_keyframe_inpaint.py:995-1024 / 741-770 |
| PDCNetAux                 | ofgen_keyframe_inpaint.py:549-653                               |
| keyframe_conv_pick        | ofgen_keyframe_inpaint.py:664-668                               |
"""
from __future__ import annotations

import glob
import os
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import ops
from .engine import RaftEngine
from .pdcnet_of import _d2h, _device, _h2d
from .pdcnet_of import warp_frame as warp_frame_pdcnet  # noqa: F401  (re-export under the scripts' alias)


# ----------------------------------------------------------------------------- W2
def warp_frame(frame: np.ndarray, flow: np.ndarray, device=None) -> np.ndarray:
    """ofgen.warp_frame: map = grid - flow, cv2.remap INTER_CUBIC, border 0.  `flow` is not modified."""
    dev = _device(device)
    out = ops.warp(_h2d(frame, dev), _h2d(np.asarray(flow, dtype=np.float32), dev), mode='cv2_cubic', sign=-1.0)
    return _d2h(out)


# ----------------------------------------------------------------------------- F0
class namespace:
    def __contains__(self, m):
        return hasattr(self, m)


class RAFT_2:
    """ofgen.py:55-79.  `model_path=None` keeps seeded random-init weights (no checkpoint ships with
    the reference); otherwise the raft-things checkpoint at `model_path` is loaded (the public files' `module.` key prefix
    is accepted, like the reference's DataParallel wrapper, ofgen.py:67-68).

    Two deliberate differences from the script's literal arithmetic, both documented in DESIGN.md §2: (1) the model runs in
    eval mode (BatchNorm running statistics, folded into the convolutions) as upstream RAFT's demo does -- the reference never
    calls `.eval()`, so its context encoder normalises with batch statistics and mutates its running stats on every call;
    (2) the all-pairs correlation runs on auto-ranged fp16 operands with fp32 accumulation (11 significant bits like the TF32
    convolutions around it, per-tensor power-of-two scaling so feature magnitude does not matter) instead of an fp32 SGEMM;
    pass `corr_precision='3xtf32'` for an fp32-faithful volume.  (3) Under torch's default `cudnn.allow_tf32 = True` -- the
    setting with which the reference itself runs TF32 convolutions on this GPU -- the encoders and the update block run cuDNN
    fp16-operand / fp32-accumulate convolutions (the same 11-bit operand significand as TF32) with the recurrent state (hidden
    state, coordinates, flow) in fp32; measured distance to the fp32 CPU reference at 768x512 / 720x1280, 20 iterations: EPE
    1.3e-3 px (tests/test_gpu_parity_full.py).  With `torch.backends.cudnn.allow_tf32 = False` every convolution is fp32."""

    def __init__(self, model_path: str | None = 'RAFT/models/raft-things.pth', iters: int = 20, device=None, **engine_kw) -> None:
        ckpt = model_path if (model_path is not None and os.path.exists(model_path)) else None
        if model_path is not None and ckpt is None:
            raise FileNotFoundError(f'{model_path} not found (pass model_path=None for random-init weights)')
        self.engine = RaftEngine(checkpoint=ckpt, iters=iters, device=device, **engine_kw)

    def to(self, device):
        self.engine.to(device)
        return self

    @torch.no_grad()
    def calc(self, img1: np.ndarray, img2: np.ndarray) -> np.ndarray:
        """BGR uint8 [H,W,3] x2 -> flow float32 [H',W',2] on img1's grid.  Like the reference the
        result keeps the replicate-padded size H',W' (multiples of 8; ofgen.py:75-78 never unpads)."""
        dev = self.engine.device
        a, b = _h2d(img1, dev)[None], _h2d(img2, dev)[None]
        return _d2h(self.engine.estimate_flow(a, b, unpad=False, bgr=True)[0])   # BGR -> RGB inside the first kernel


def create_of_algo(model_path: str | None = 'RAFT/models/raft-things.pth', **kw):
    """ofgen.py:81-83."""
    return RAFT_2(model_path, **kw)


def of_calc(frame1: np.ndarray, frame2: np.ndarray, of_algo):
    """ofgen.py:45-49: (flow, |flow|)."""
    flow = of_algo.calc(frame1, frame2)
    fx, fy = flow[:, :, 0], flow[:, :, 1]
    return flow, np.sqrt(fx * fx + fy * fy)


def of_calc_pdcnet(frame1: np.ndarray, frame2: np.ndarray, algo, device=None):
    """ofgen_pixel_inpaint.py:105-118: (flow, confidence, travel distance, log_confidence)."""
    flow, confidence, log_confidence = algo.calc(frame1, frame2)
    dev = _device(device)
    v = _d2h(ops.travel_distance(_h2d(flow, dev)[None], _h2d(confidence, dev)[None], 0.9)[0])
    return flow, confidence, v, log_confidence


# ----------------------------------------------------------------------------- masks
def generate_mask(cum_confidence: np.ndarray, log_confidence: np.ndarray, thres: float = 0.8, device=None):
    """Returns (mask u8, log_confidence); like the reference, `log_confidence` is also reset IN PLACE
    where the confidence is below `thres`."""
    dev = _device(device)
    conf_d = _h2d(np.asarray(cum_confidence, dtype=np.float32), dev)[None]
    logc_d = _h2d(np.asarray(log_confidence, dtype=np.float32), dev)[None].clone()
    mask = _d2h(ops.generate_mask(conf_d, logc_d, thres, 7)[0])
    log_confidence[...] = logc_d[0].cpu().numpy()
    return mask, log_confidence


def confidence_to_mask(confidence, flow, dist, mask_aux, device=None):
    """ofgen_pixel_inpaint.py:218-227; `mask_aux` carries .pixel_travel_dist and .thres and is updated."""
    dev = _device(device)
    conf_d = _h2d(np.asarray(confidence, dtype=np.float32), dev)
    low = conf_d < 0.9
    ptd = ops.warp(_h2d(np.asarray(mask_aux.pixel_travel_dist, dtype=np.float32), dev),
                   _h2d(np.asarray(flow, dtype=np.float32), dev), mode='cv2_cubic', sign=1.0)
    ptd = ptd + _h2d(np.asarray(dist, dtype=np.float32), dev)
    ptd[low] = 0
    far = ptd > mask_aux.thres
    mask = torch.zeros_like(conf_d, dtype=torch.uint8)
    mask[low | far] = 255
    ptd[far] = 0
    mask_aux.pixel_travel_dist = ptd.cpu().numpy()
    return _d2h(ops.dilate_ellipse(mask[None].contiguous(), 15)[0])


def mix_propagated_ai_frame(raw_ai_frame, warped_propagated_ai_frame, mask, propagated_pixel_weight=1.0, device=None):
    if propagated_pixel_weight < 0.001:
        return raw_ai_frame
    dev = _device(device)
    out = ops.mix_propagated(_h2d(raw_ai_frame, dev)[None], _h2d(warped_propagated_ai_frame, dev)[None],
                             _h2d(mask, dev)[None], propagated_pixel_weight)
    return _d2h(out[0])


def merge_images(base_image, second_image, mask, method='naive', device=None):
    if method != 'naive':
        raise NotImplementedError("only method='naive' is on the hot path (the 'poisson' branch calls cv2.seamlessClone)")
    dev = _device(device)
    out = ops.merge_select(_h2d(base_image, dev)[None], _h2d(second_image, dev)[None], _h2d(mask, dev)[None])
    return _d2h(out[0])


def expand_mask(mask: np.ndarray, ori_image: np.ndarray, device=None) -> np.ndarray:
    dev = _device(device)
    return _d2h(ops.expand_mask(_h2d(mask, dev)[None], _h2d(ori_image, dev)[None], 7)[0])


def invert_and_dilate(mask: np.ndarray, device=None) -> np.ndarray:
    """mask2 = dilate(255 - mask, ellipse 7x7) (ofgen_keyframe_inpaint.py:772-774)."""
    dev = _device(device)
    return _d2h(ops.dilate_ellipse(_h2d(mask, dev)[None], 7, invert=True)[0])


def composite_references(flow_mat: np.ndarray, ai_frames, thres: float = 0.5, device=None):
    """The greedy multi-reference warp+composite loop (ofgen_keyframe_inpaint.py:995-1024).
    flow_mat [n,1,H,W,3] is updated in place exactly like the reference; ai_frames: n uint8 [H,W,3].
    Returns (ret_frame u8 [H,W,3], mask u8 [H,W], chosen reference order)."""
    dev = _device(device)
    n, one, H, W, _ = flow_mat.shape
    assert one == 1
    fm = _h2d(flow_mat.reshape(n, H, W, 3).astype(np.float32, copy=False), dev).clone()
    frames = _h2d(np.stack([np.asarray(f) for f in ai_frames]), dev)
    ret, mask, order = ops.greedy_composite(fm, frames, thres)
    flow_mat[...] = fm.cpu().numpy().reshape(flow_mat.shape)
    return _d2h(ret), _d2h(mask), [int(i) for i in order.cpu().tolist()]


# ----------------------------------------------------------------------------- A1 / A2
def chunks(lst, n):
    for i in range(0, len(lst), n):
        yield lst[i:i + n]


class PDCNetAux:
    """Batched pair flow with the on-disk `.npy` pair cache (ofgen_keyframe_inpaint.py:549-653).
    `video` needs `.size_hw` and `.get_raw_frame(i) -> BGR uint8 [H,W,3]`; `indices` arguments are
    plain lists of frame numbers (or objects with `.indices`)."""

    def __init__(self, pdcnet_model, workspace_dir: str, batch_size: int = 16, device=None) -> None:
        self.workspace_dir = workspace_dir
        self.cached_pair = set()
        self.batch_size = batch_size
        self.device = _device(device)
        self.pdcnet_model = pdcnet_model.to(self.device)
        self.pair_dir = os.path.join(workspace_dir, 'pdcnet')
        os.makedirs(self.pair_dir, exist_ok=True)
        for f in glob.glob(os.path.join(self.pair_dir, '*.npy')):
            s, t = os.path.split(f)[-1].split('.')[0].split('-')
            self.cached_pair.add((int(s), int(t)))

    def purge(self):
        self.cached_pair = set()
        for f in glob.glob(os.path.join(self.pair_dir, '*.npy')):
            os.remove(f)

    def load_cached(self, s, t):
        assert (s, t) in self.cached_pair
        return np.load(os.path.join(self.pair_dir, f'{s:05d}-{t:05d}.npy'))

    def calcualte_single(self, video, s, t):  # (sic) the reference's spelling, ofgen_keyframe_inpaint.py:576
        if (s, t) in self.cached_pair:
            return self.load_cached(s, t)
        ret = np.zeros((1, 1, *video.size_hw, 3), dtype=np.float32)
        self.calculate_given_pairs(video, [(s, t)], {s: 0}, {t: 0}, ret)
        self.cached_pair.add((s, t))
        return ret[0, 0]

    calculate_single = calcualte_single

    def calculate_given_pairs(self, video, to_calculate_pairs: List[Tuple[int, int]], s2i_map: Dict[int, int],
                              t2i_map: Dict[int, int], ret: np.ndarray):
        for pair_batch in chunks(to_calculate_pairs, self.batch_size):
            src = np.stack([video.get_raw_frame(s)[:, :, ::-1] for s, _ in pair_batch])  # BGR -> RGB
            tgt = np.stack([video.get_raw_frame(t)[:, :, ::-1] for _, t in pair_batch])
            flow_est, confidence = self.pdcnet_model.calc_batch(_h2d(src, self.device), _h2d(tgt, self.device))
            for i, (s, t) in enumerate(pair_batch):
                si, ti = s2i_map[s], t2i_map[t]
                ret[si, ti, :, :, 0:2] = flow_est[i]
                ret[si, ti, :, :, 2] = confidence[i]
                np.save(os.path.join(self.pair_dir, f'{s:05d}-{t:05d}.npy'), ret[si, ti])

    @staticmethod
    def _idx(indices):
        return list(indices.indices) if hasattr(indices, 'indices') else list(indices)

    def calculate_multiple_to_one(self, video, source_indices, target_index: int) -> np.ndarray:
        """[n_sources, 1, H, W, 3] (flow x, flow y, confidence); identity pair = zero flow, confidence 1."""
        srcs = self._idx(source_indices)
        s2i = {s: i for i, s in enumerate(srcs)}
        todo = [(s, target_index) for s in srcs if s != target_index and (s, target_index) not in self.cached_pair]
        ret = np.zeros((len(srcs), 1, *video.size_hw, 3), dtype=np.float32)
        self.calculate_given_pairs(video, todo, s2i, {target_index: 0}, ret)
        for i, s in enumerate(srcs):
            if s == target_index:
                ret[i, 0, :, :, 0:2] = 0
                ret[i, 0, :, :, 2] = 1
            elif (s, target_index) in self.cached_pair:
                ret[i, 0] = self.load_cached(s, target_index)
        self.cached_pair.update(todo)
        return ret

    @torch.no_grad()
    def calculate_multiple_to_one_device(self, source_frames_rgb: torch.Tensor, target_frame_rgb: torch.Tensor,
                                         identity: List[bool] | None = None) -> torch.Tensor:
        """Device-resident form of `calculate_multiple_to_one` for callers that already hold the frames on the GPU: RGB uint8
        sources [n,H,W,3] and one target [H,W,3] (CUDA) -> flow_mat [n,1,H,W,3] fp32 on the device (flow x, flow y, confidence),
        the layout `composite_references` / `ops.greedy_composite` consume -- no PNG decode, H2D, D2H or `.npy` file per pair
        (ofgen_keyframe_inpaint.py:585-625 does all four).  `identity[i]` marks a source that IS the target (flow 0,
        confidence 1, :621-623).  Pairs run through `calc_batch_device` in chunks of `batch_size`."""
        n, H, W, _ = source_frames_rgb.shape
        dev = self.device
        out = torch.zeros((n, 1, H, W, 3), dtype=torch.float32, device=dev)
        todo = [i for i in range(n) if not (identity and identity[i])]
        tgt1 = target_frame_rgb.to(dev)[None]
        for chunk in chunks(todo, self.batch_size):
            idx = torch.tensor(chunk, device=dev)
            src = source_frames_rgb.to(dev).index_select(0, idx).contiguous()
            flow, conf = self.pdcnet_model.calc_batch_device(src, tgt1.expand(len(chunk), -1, -1, -1).contiguous())
            out[idx, 0, :, :, 0:2] = flow
            out[idx, 0, :, :, 2] = conf
        for i in range(n):
            if identity and identity[i]:
                out[i, 0, :, :, 2] = 1.0
        return out

    def calculate_pairwise(self, video, indices) -> np.ndarray:
        """[n, n, H, W, 3] for every ordered pair of `indices`."""
        idx = self._idx(indices)
        s2i = {s: i for i, s in enumerate(idx)}
        todo = [(s, t) for s in idx for t in idx if s != t and (s, t) not in self.cached_pair]
        ret = np.zeros((len(idx), len(idx), *video.size_hw, 3), dtype=np.float32)
        self.calculate_given_pairs(video, todo, s2i, dict(s2i), ret)
        for i, s in enumerate(idx):
            for j, t in enumerate(idx):
                if s == t:
                    ret[i, j, :, :, 0:2] = 0
                    ret[i, j, :, :, 2] = 1
                elif (s, t) in self.cached_pair:
                    ret[i, j] = self.load_cached(s, t)
        self.cached_pair.update(todo)
        return ret


def keyframe_conv_pick(flow_mat: np.ndarray, device=None) -> int:
    """argmax_s sum_{t,h,w} confidence[s,t] (ofgen_keyframe_inpaint.py:664-668), reduced on the device."""
    dev = _device(device)
    sums = ops.confidence_sums(_h2d(np.asarray(flow_mat, dtype=np.float32), dev))
    return int(torch.argmax(sums).item())


# ----------------------------------------------------------------------------- key-frame detector (before the path)
def estimated_kernel_size(frame_width: int, frame_height: int) -> int:
    """ofgen_pixel_inpaint.py:142-147."""
    import math
    size = 4 + round(math.sqrt(frame_width * frame_height) / 192)
    if size % 2 == 0:
        size += 1
    return size


def detect_edges_device(frame_bgr: torch.Tensor) -> torch.Tensor:
    """detect_edges for a CUDA uint8 [H,W,3] frame -> CUDA uint8 [H,W] edge map (stays on the device)."""
    H, W, _ = frame_bgr.shape
    return ops.detect_edges(frame_bgr, estimated_kernel_size(W, H))


def detect_edges(frame: np.ndarray, device=None) -> np.ndarray:
    """ofgen_pixel_inpaint.py:150-176: HSV value channel -> median-thresholded Canny -> k x k dilation."""
    return _d2h(detect_edges_device(_h2d(frame, _device(device))))


def mean_pixel_distance(left, right, device=None) -> float:
    """ofgen_pixel_inpaint.py:132-139 (from PySceneDetect): mean |left - right| of two 2-D 8-bit images."""
    if isinstance(left, np.ndarray):
        dev = _device(device)
        left, right = _h2d(left, dev), _h2d(right, dev)
    assert left.dim() == 2 and right.dim() == 2
    assert left.shape == right.shape
    return ops.abs_diff_sum(left.contiguous(), right.contiguous()) / float(left.shape[0] * left.shape[1])


class KeyFrameSelector:
    """The decision logic of frame_generator (ofgen_pixel_inpaint.py:272-313) without the video IO: feed the resized
    frames in order, `push(frame)` answers whether the frame starts a new key-frame segment.  Edge maps stay on the
    device; only the scalar distance comes back."""

    def __init__(self, fps: float = 30.0, th: float = 8.5, min_gap: int = -1, max_gap: int = -1, device=None):
        self.th = th
        self.min_gap = int(10 * fps / 30) if min_gap == -1 else int(max(1, min_gap) * fps / 30)
        self.max_gap = int(300 * fps / 30) if max_gap == -1 else int(max(10, max_gap) * fps / 30)
        self.gap = 0
        self.key_edges = None
        self.device = _device(device)

    def push(self, frame, frames_advanced: int = 1) -> bool:
        """frame: BGR uint8 [H,W,3] (numpy or CUDA tensor); frames_advanced = source frames consumed since the last
        push (`keep_every`), which is what the reference's `gap` counts (:293-297)."""
        self.gap += frames_advanced
        f = frame if isinstance(frame, torch.Tensor) else _h2d(frame, self.device)
        edges = detect_edges_device(f)
        if self.key_edges is None:
            self.key_edges = edges
            return True
        delta = mean_pixel_distance(edges, self.key_edges)
        if self.th * (self.max_gap - self.gap) / self.max_gap < delta:
            self.key_edges = edges
            self.gap = 0
            return True
        return False
