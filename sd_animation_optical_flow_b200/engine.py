"""Device-resident flow engine: the batched, stream-ordered form of the reference's
`RAFT_2.calc` (ofgen.py:55-79) plus the `estimate_flow()` / `warp()` call surface the
north star names, so a clip's flow -> warp -> mask -> composite never leaves the GPU.

    eng  = RaftEngine(iters=20)                       # random-init or eng.load_checkpoint(path)
    flow = eng.estimate_flow(img1_u8, img2_u8)        # [B,H,W,3] RGB uint8 CUDA -> [B,H,W,2] fp32
    out  = warp(stylised_key_u8, flow)                # bit-exact cv2.remap cubic on the device

The convolutions run in PyTorch/cuDNN; the all-pairs correlation volume + pyramid, the
per-iteration lookup, the warp and the mask/composite steps run in this package's sm_100a
kernels.  The whole forward for one input shape can be captured into a CUDA graph
(`use_cuda_graph`, the default on the fast path) so the ~480 launches of a 20-iteration pass replay without host work.
"""
from __future__ import annotations

import contextlib
from collections import OrderedDict
from types import SimpleNamespace

import torch

from . import ops
from .raft import RAFT, InputPadder, fill_weights_by_name


def load_raft_state_dict(path: str) -> dict:
    """State dict of a RAFT checkpoint.  The public `raft-*.pth` files were saved from `nn.DataParallel(RAFT(args))`
    (which is why the reference wraps its model the same way, ofgen.py:67-68), so their keys carry a `module.` prefix."""
    sd = torch.load(path, map_location='cpu')
    if isinstance(sd, dict) and 'state_dict' in sd and not any(k.endswith('.weight') for k in sd):
        sd = sd['state_dict']
    return {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in sd.items()}


def warp(img: torch.Tensor, flow: torch.Tensor, mode: str = 'cv2_cubic', sign: float = 1.0) -> torch.Tensor:
    """Backward warp of CUDA tensors (see ops.warp)."""
    return ops.warp(img, flow, mode=mode, sign=sign)


class RaftEngine:
    def __init__(self, checkpoint: str | None = None, iters: int = 20, small: bool = False,
                 corr_precision: str = 'fp16', alternate_corr: bool = False, mixed_precision: bool = False,
                 channels_last: bool = False, use_cuda_graph: bool | None = None, device=None, seed: int = 0,
                 fast: bool | None = None, cudnn_benchmark: bool = True, fast_options: dict | None = None,
                 max_graphs: int = 4, flow_head_scale: float = 1.0):
        if not torch.cuda.is_available():
            raise RuntimeError('RaftEngine needs a CUDA device (B200); there is no CPU path')
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.iters = iters
        # let cuDNN time its engines once per convolution shape (during the first call / the graph warm-up): its heuristics
        # pick slow kernels for several of RAFT's small-batch shapes.  Scoped to this engine's own convolutions
        # (`_cudnn_scope`), not set process-wide: the host process also runs the Stable-Diffusion model.
        self.cudnn_benchmark = bool(cudnn_benchmark)
        self.args = SimpleNamespace(small=small, mixed_precision=mixed_precision, alternate_corr=alternate_corr,
                                    corr_precision=corr_precision)
        model = RAFT(self.args)
        if checkpoint is None:
            fill_weights_by_name(model, seed, flow_head_scale)   # no checkpoint: name-seeded random-init weights
        else:
            model.load_state_dict(load_raft_state_dict(checkpoint))
        model = model.to(self.device).eval()
        if channels_last:
            model = model.to(memory_format=torch.channels_last)
        self.model = model
        self.channels_last = channels_last
        self._graphs = OrderedDict()      # (kind, input shape, iters, bgr) -> (graph, static inputs, static output), LRU
        self.max_graphs = max_graphs
        # the hand-scheduled NHWC forward (raft_fast.py) is the default for the configuration the scripts use
        if fast is None:
            fast = not small and not alternate_corr and not mixed_precision and not channels_last
        self.fast = None
        self.fast_options = dict(fast_options or {})
        if fast:
            from .raft_fast import FastRaft
            self.fast = FastRaft(self.model, corr_precision, **self.fast_options)
        # CUDA graphs by default on the fast path (7.2 -> 4.0 ms per 768x512 pair): one graph per input shape, captured at
        # the first call with that shape (about a second), the `max_graphs` most recently used shapes are kept
        self.use_cuda_graph = (self.fast is not None) if use_cuda_graph is None else bool(use_cuda_graph)

    def load_checkpoint(self, path: str):
        self.model.load_state_dict(load_raft_state_dict(path))
        self._graphs.clear()
        if self.fast is not None:
            from .raft_fast import FastRaft
            self.fast = FastRaft(self.model, self.args.corr_precision, **self.fast_options)
        return self

    def to(self, device):
        self.device = torch.device(device)
        self.model = self.model.to(self.device)
        self._graphs.clear()
        if self.fast is not None:
            from .raft_fast import FastRaft
            self.fast = FastRaft(self.model, self.args.corr_precision, **self.fast_options)
        return self

    @contextlib.contextmanager
    def _scope(self):
        """Everything this engine launches runs with ITS device current (the C library launches on the current device and
        keys its tables on it) and, on request, with cuDNN's benchmark mode on for this engine's convolutions only."""
        prev = torch.backends.cudnn.benchmark
        try:
            if self.cudnn_benchmark:
                torch.backends.cudnn.benchmark = True
            with torch.cuda.device(self.device):
                yield
        finally:
            torch.backends.cudnn.benchmark = prev

    def _capture_stream(self):
        """Capture stream on THIS engine's device.  torch.cuda.graph's default capture stream is a process-wide singleton
        created on whichever device captured first; reusing it for an engine on another device switches the current device
        inside the capture and invalidates it."""
        st = getattr(self, '_cap_stream', None)
        if st is None or st.device != self.device:
            st = self._cap_stream = torch.cuda.Stream(device=self.device)
        return st

    def _graph_get(self, key):
        ent = self._graphs.get(key)
        if ent is not None:
            self._graphs.move_to_end(key)
        return ent

    def _graph_put(self, key, ent):
        self._graphs[key] = ent
        while len(self._graphs) > max(1, self.max_graphs):
            self._graphs.popitem(last=False)   # least recently used shape: its graph and static buffers are released

    # ---------------------------------------------------------------- core
    @torch.no_grad()
    def _forward(self, im1: torch.Tensor, im2: torch.Tensor) -> torch.Tensor:
        """padded float images [B,3,H,W] in 0..255 -> flow_up [B,2,H,W]."""
        if self.fast is not None:
            _, flow_up = self.fast.forward(im1, im2, self.iters)
            return flow_up.permute(0, 3, 1, 2)
        if self.channels_last:
            im1 = im1.contiguous(memory_format=torch.channels_last)
            im2 = im2.contiguous(memory_format=torch.channels_last)
        _, flow_up = self.model(im1, im2, iters=self.iters, test_mode=True)
        return flow_up.float()

    @torch.no_grad()
    def _forward_graphed(self, im1: torch.Tensor, im2: torch.Tensor) -> torch.Tensor:
        key = (tuple(im1.shape), self.iters)
        ent = self._graph_get(key)
        if ent is None:
            s1, s2 = im1.clone(), im2.clone()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up: cuDNN autotune, table uploads, allocator
                    self._forward(s1, s2)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._capture_stream()):
                out = self._forward(s1, s2)
            ent = (g, s1, s2, out)
            self._graph_put(key, ent)
        g, s1, s2, out = ent
        s1.copy_(im1)
        s2.copy_(im2)
        g.replay()
        return out.clone()  # the graph's output buffer is overwritten by the next replay

    # ---------------------------------------------------------------- uint8 fast path
    @torch.no_grad()
    def _forward_u8(self, a: torch.Tensor, b: torch.Tensor, pad, bgr: bool = False) -> torch.Tensor:
        """uint8 RGB [B,H,W,3] x2 -> padded-size flow [B,Hp,Wp,2]: one kernel normalises + pads each frame straight into
        the channels-last layout the encoders consume (csrc/raft_glue.cu::normalize_pad_u8_nhwc_kernel)."""
        im1 = ops.normalize_pad_u8(a, pad, channels=4, bgr=bgr)
        im2 = ops.normalize_pad_u8(b, pad, channels=4, bgr=bgr)
        _, flow_up = self.fast.forward(im1, im2, self.iters, normalized=True)
        return flow_up

    @torch.no_grad()
    def _forward_u8_graphed(self, a: torch.Tensor, b: torch.Tensor, pad, bgr: bool = False) -> torch.Tensor:
        key = ('u8', tuple(a.shape), self.iters, bool(bgr))
        ent = self._graph_get(key)
        if ent is None:
            s1, s2 = a.clone(), b.clone()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up: cuDNN autotune, table uploads, allocator
                    self._forward_u8(s1, s2, pad, bgr)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._capture_stream()):
                out = self._forward_u8(s1, s2, pad, bgr)
            ent = (g, s1, s2, out)
            self._graph_put(key, ent)
        g, s1, s2, out = ent
        s1.copy_(a)
        s2.copy_(b)
        g.replay()
        return out  # overwritten by the next replay: estimate_flow copies (unpad / contiguous) before returning

    # ---------------------------------------------------------------- consecutive frames of a clip
    def _forward_seq(self, frames: torch.Tensor, pad, bgr: bool) -> torch.Tensor:
        im = ops.normalize_pad_u8(frames, pad, channels=4, bgr=bgr)
        _, flow_up = self.fast.forward(im, None, self.iters, normalized=True, sequence=True)
        return flow_up

    @torch.no_grad()
    def estimate_flow_sequence(self, frames: torch.Tensor, unpad: bool = True, bgr: bool = False) -> torch.Tensor:
        """Flows of the consecutive pairs of a clip: frames [n,H,W,3] (uint8 CUDA) -> [n-1,H,W,2] fp32, pair i = RAFT(image1 =
        frame i, image2 = frame i+1) -- the loop of ofgen.py (`calc(prev, cur)` for every frame).  Equal to
        estimate_flow(frames[:-1], frames[1:]) but the feature encoder runs once per frame instead of twice (frame i+1 is
        image2 of pair i and image1 of pair i+1).  One CUDA graph per shape."""
        if self.fast is None:
            raise RuntimeError('estimate_flow_sequence needs the fast path (basic model, fast=True)')
        if frames.dim() != 4 or frames.shape[-1] != 3 or frames.dtype != torch.uint8 or not frames.is_cuda or frames.shape[0] < 2:
            raise RuntimeError(f'frames must be a uint8 CUDA tensor [n >= 2,H,W,3], got {tuple(frames.shape)} {frames.dtype}')
        with self._scope():
            n, H, W, _ = frames.shape
            pad = InputPadder((H, W))._pad
            frames = frames.contiguous()
            if self.use_cuda_graph:
                gk = ('seq', tuple(frames.shape), self.iters, bool(bgr))
                ent = self._graph_get(gk)
                if ent is None:
                    s1 = frames.clone()
                    side = torch.cuda.Stream(device=self.device)
                    side.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(side):
                        for _ in range(2):
                            self._forward_seq(s1, pad, bgr)
                    torch.cuda.current_stream(self.device).wait_stream(side)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self._capture_stream()):
                        out = self._forward_seq(s1, pad, bgr)
                    ent = (g, s1, out)
                    self._graph_put(gk, ent)
                g, s1, out = ent
                s1.copy_(frames)
                g.replay()
                flow_up = out
            else:
                flow_up = self._forward_seq(frames, pad, bgr)
            if unpad and any(pad):
                Hp, Wp = flow_up.shape[1:3]
                flow_up = flow_up[:, pad[2]:Hp - pad[3], pad[0]:Wp - pad[1]]
            return flow_up.clone(memory_format=torch.contiguous_format) if self.use_cuda_graph else flow_up.contiguous()

    # ---------------------------------------------------------------- key-frame scheme
    @torch.no_grad()
    def encode_key(self, key_img: torch.Tensor, bgr: bool = False):
        """Encode ONE key frame (uint8 [H,W,3] or [1,H,W,3], CUDA) for `estimate_flow_keyed`: fnet(key) and the pooled 16-bit
        correlation operands are computed here, once, instead of once per pair (ofgen_pixel_inpaint.py:335 and
        ofgen_keyframe_inpaint.py:602-625 run every non-key frame against the same key / reference frames)."""
        if self.fast is None:
            raise RuntimeError('key-frame feature reuse needs the fast path (basic model, fast=True)')
        if key_img.dim() == 3:
            key_img = key_img[None]
        if key_img.dim() != 4 or key_img.shape[0] != 1 or key_img.shape[-1] != 3 or key_img.dtype != torch.uint8 or not key_img.is_cuda:
            raise RuntimeError(f'key frame must be a uint8 CUDA tensor [H,W,3] or [1,H,W,3], got {tuple(key_img.shape)} {key_img.dtype}')
        with self._scope():
            H, W = key_img.shape[1:3]
            pad = InputPadder((H, W))._pad
            im = ops.normalize_pad_u8(key_img.contiguous(), pad, channels=4, bgr=bgr)
            kf = self.fast.encode_key(im, normalized=True)
            kf.shape, kf.bgr = (H, W), bool(bgr)
            return kf

    @torch.no_grad()
    def _forward_keyed(self, frames: torch.Tensor, key, pad, bgr: bool) -> torch.Tensor:
        im1 = ops.normalize_pad_u8(frames, pad, channels=4, bgr=bgr)
        _, flow_up = self.fast.forward(im1, None, self.iters, normalized=True, key=key)
        return flow_up

    @torch.no_grad()
    def estimate_flow_keyed(self, key, frames: torch.Tensor, unpad: bool = True, bgr: bool | None = None) -> torch.Tensor:
        """Flow of every frame [B,H,W,3] (uint8 CUDA) -> the key frame, on the FRAME's grid: RAFT(image1 = frame, image2 = key),
        i.e. `frame(x) ~ key(x + flow(x))`, the convention `pdcnet_of.warp_frame(key_ai, flow)` consumes (x + flow) and the
        one in which the key's features are shared by all pairs.  `key` = encode_key(key_img).  Returns [B,H,W,2] fp32.
        With CUDA graphs the graph is captured per (shape, key object): replays for the same key reuse its operands."""
        if self.fast is None:
            raise RuntimeError('key-frame feature reuse needs the fast path (basic model, fast=True)')
        if frames.dim() != 4 or frames.shape[-1] != 3 or frames.dtype != torch.uint8 or not frames.is_cuda:
            raise RuntimeError(f'frames must be a uint8 CUDA tensor [B,H,W,3], got {tuple(frames.shape)} {frames.dtype}')
        if tuple(frames.shape[1:3]) != tuple(key.shape):
            raise RuntimeError(f'frames are {tuple(frames.shape[1:3])}, the key frame was {tuple(key.shape)}')
        bgr = key.bgr if bgr is None else bool(bgr)
        with self._scope():
            B, H, W, _ = frames.shape
            pad = InputPadder((H, W))._pad
            frames = frames.contiguous()
            if self.use_cuda_graph:
                # the key's tensors are baked into the graph as addresses: one graph per shape, re-pointed at a new key by
                # copying the new key's feature map / operand buffers into the static ones
                gk = ('keyed', tuple(frames.shape), self.iters, bgr)
                ent = self._graph_get(gk)
                if ent is None:
                    s1 = frames.clone()
                    skey = type(key)(key.fmap2.clone(), None)
                    if key.target is not None:
                        skey.target = ops.CorrTarget.__new__(ops.CorrTarget)
                        skey.target.__dict__.update(key.target.__dict__)
                        skey.target.buf = key.target.buf.clone()
                    side = torch.cuda.Stream(device=self.device)
                    side.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(side):
                        for _ in range(2):
                            self._forward_keyed(s1, skey, pad, bgr)
                    torch.cuda.current_stream(self.device).wait_stream(side)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self._capture_stream()):
                        out = self._forward_keyed(s1, skey, pad, bgr)
                    ent = (g, s1, skey, out, [None])
                    self._graph_put(gk, ent)
                g, s1, skey, out, cur = ent
                if cur[0] is not key:            # a different key than the one the static buffers hold: refresh them
                    skey.fmap2.copy_(key.fmap2)
                    if key.target is not None:
                        skey.target.buf.copy_(key.target.buf)
                    cur[0] = key
                s1.copy_(frames)
                g.replay()
                flow_up = out
            else:
                flow_up = self._forward_keyed(frames, key, pad, bgr)
            if unpad and any(pad):
                Hp, Wp = flow_up.shape[1:3]
                flow_up = flow_up[:, pad[2]:Hp - pad[3], pad[0]:Wp - pad[1]]
            return flow_up.clone(memory_format=torch.contiguous_format) if self.use_cuda_graph else flow_up.contiguous()

    @torch.no_grad()
    def estimate_flow(self, img1: torch.Tensor, img2: torch.Tensor, unpad: bool = True, bgr: bool = False) -> torch.Tensor:
        """img1, img2: RGB (bgr=True: BGR) uint8 (or float 0..255) CUDA tensors [B,H,W,3].  Returns flow of img1 -> img2
        on img1's grid, [B,H,W,2] fp32 (x, y pixels).  Images are replicate-padded to a multiple of 8
        like the reference (utils/utils.py:7-19); unpad=False reproduces RAFT_2.calc, which returns
        the padded-size flow (ofgen.py:75-78).  The result is always a fresh tensor owned by the caller."""
        if img1.dim() != 4 or img1.shape[-1] != 3 or img1.shape != img2.shape:
            raise RuntimeError(f'images must both be [B,H,W,3], got {tuple(img1.shape)} and {tuple(img2.shape)}')
        if not img1.is_cuda or not img2.is_cuda:
            raise RuntimeError('estimate_flow takes CUDA tensors; use ofgen.RAFT_2.calc for numpy frames')
        if img1.device != self.device or img2.device != self.device:
            raise RuntimeError(f'images must live on the engine\'s device {self.device}, got {img1.device} / {img2.device}')
        with self._scope():
            if self.fast is not None and img1.dtype == torch.uint8 and img2.dtype == torch.uint8:
                B, H, W, _ = img1.shape
                pad = InputPadder((H, W))._pad
                fwd = self._forward_u8_graphed if self.use_cuda_graph else self._forward_u8
                flow_up = fwd(img1.contiguous(), img2.contiguous(), pad, bgr)
                if unpad and any(pad):
                    Hp, Wp = flow_up.shape[1:3]
                    flow_up = flow_up[:, pad[2]:Hp - pad[3], pad[0]:Wp - pad[1]]
                if self.use_cuda_graph:
                    # the graph's static output buffer is overwritten by the next replay: ALWAYS hand out a copy.  (A slice of
                    # a B=1 tensor that only trims rows is already "contiguous", so .contiguous() would return the view.)
                    return flow_up.clone(memory_format=torch.contiguous_format)
                return flow_up.contiguous()
            if bgr:
                img1, img2 = img1.flip(-1), img2.flip(-1)
            im1 = img1.permute(0, 3, 1, 2).float()
            im2 = img2.permute(0, 3, 1, 2).float()
            padder = InputPadder(im1.shape)
            im1, im2 = padder.pad(im1, im2)
            fwd = self._forward_graphed if self.use_cuda_graph else self._forward
            flow_up = fwd(im1.contiguous(), im2.contiguous())
            if unpad:
                flow_up = padder.unpad(flow_up)
            return flow_up.permute(0, 2, 3, 1).contiguous()


class RaftFlowConfidence:
    """Adapter giving a RAFT engine the DenseMatching interface the reference's PDCNetPlus
    wraps (`estimate_flow_and_confidence_map(source, target)`, pdcnet_of.py:70):

      * flow lives on the TARGET grid and points into the SOURCE (so `warp_frame(source_ai, flow)`
        samples at x + flow, pdcnet_of.py:39-41): it is RAFT(image1=target, image2=source);
      * RAFT has no confidence head, so `weight_map` is defined HERE (nothing in the reference to
        match, SURVEY §8a note): forward-backward consistency,
            err  = |f_ts(x) + f_st(x + f_ts(x))|              (bilinear warp kernel)
            weight_map = stack([bias - err / sigma, 0])  =>  confidence = sigmoid(bias - err/sigma).
    """

    def __init__(self, engine: RaftEngine, sigma: float = 1.0, bias: float = 4.0, synthetic_logits_seed: int | None = None):
        self.engine = engine
        self.sigma, self.bias = sigma, bias
        self.synthetic_logits_seed = synthetic_logits_seed

    def to(self, device):
        self.engine.to(device)
        return self

    @torch.no_grad()
    def estimate_flow_and_confidence_map(self, source: torch.Tensor, target: torch.Tensor):
        """source, target: [B,3,H,W] uint8/float RGB.  Returns (flow [B,2,H,W], {'weight_map': [B,2,H,W]})."""
        dev = self.engine.device
        src = source.to(dev).permute(0, 2, 3, 1).contiguous()
        tgt = target.to(dev).permute(0, 2, 3, 1).contiguous()
        f_ts = self.engine.estimate_flow(tgt, src)  # on the target grid, into the source
        if self.synthetic_logits_seed is not None:
            g = torch.Generator(device=dev).manual_seed(self.synthetic_logits_seed)
            B, H, W, _ = f_ts.shape
            wm = 2.0 * torch.randn((B, 2, H, W), generator=g, device=dev)
        else:
            f_st = self.engine.estimate_flow(src, tgt)  # on the source grid, into the target
            back = ops.warp(f_st, f_ts, mode='bilinear', sign=1.0)  # f_st sampled at x + f_ts(x)
            err = torch.linalg.vector_norm(f_ts + back, dim=-1)
            wm = torch.stack([self.bias - err / self.sigma, torch.zeros_like(err)], dim=1)
        return f_ts.permute(0, 3, 1, 2).contiguous(), {'weight_map': wm.contiguous()}


_default_engine: RaftEngine | None = None


def estimate_flow(img1: torch.Tensor, img2: torch.Tensor, iters: int | None = None, engine: RaftEngine | None = None) -> torch.Tensor:
    """Module-level convenience: flow img1 -> img2 for [B,H,W,3] CUDA images.  `engine` = a caller-configured engine (its
    own `iters` is used; passing a different `iters` together with an engine is an error, not a silent reconfiguration);
    without one a process-wide default engine is created with `iters` (20 = ofgen.py:77 when omitted)."""
    global _default_engine
    if engine is not None:
        if iters is not None and iters != engine.iters:
            raise ValueError(f'iters={iters} conflicts with the supplied engine (iters={engine.iters}); configure the engine instead')
        return engine.estimate_flow(img1, img2)
    want = 20 if iters is None else int(iters)
    if _default_engine is None or _default_engine.iters != want or _default_engine.device != img1.device:
        _default_engine = RaftEngine(iters=want, device=img1.device)
    return _default_engine.estimate_flow(img1, img2)
