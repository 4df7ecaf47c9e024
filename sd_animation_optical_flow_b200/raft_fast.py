"""Hand-scheduled channels-last RAFT forward (inference, basic model): the same arithmetic as
`raft.RAFT.forward(test_mode=True)` / the reference's RAFT/core/raft.py:86-144, re-plumbed so that one GRU
iteration is ~25 kernels instead of ~100:

  * every activation of the update block lives in a dense NHWC buffer, so cuDNN's tensor-core convolutions
    run without the NCHW<->NHWC transposes eager PyTorch wraps around each of them;
  * the `torch.cat`s of update.py (hx = [h, x], [r*h, x], [cor, flo], [out, flow]) are persistent
    concatenated buffer HX = [h | motion | flow] whose channel slices are written in place by the producers
    (the other concatenations of update.py disappear algebraically, see FastRaft.__init__);
  * convz and convr (same input) are one convolution with concatenated filters;
  * the element-wise work between convolutions is four hand-written kernels (csrc/raft_glue.cu), the
    correlation lookup writes channels-last directly (csrc/corr_lookup.cu), and the convex 8x upsample
    is one kernel producing the [B,H,W,2] flow the warp kernel consumes;
  * the two convolutions cuDNN serves badly (7x7 on the 2-channel flow; 3x3 down to 2 channels) are
    hand-written fp32 kernels, the second fused with the coords/flow update;
  * the encoders run channels-last as well (FastEncoder), the context encoder on a side stream next to
    the feature encoder, and the flow branch of the motion encoder next to the correlation branch.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .corr import _to_nhwc

CL = torch.channels_last


def _warn_fallback(who: str, why: str) -> None:
    """The fused bias + ReLU convolution is checked once against the plain one; falling back is a performance cliff
    (one more kernel per convolution), so it is never silent."""
    import warnings
    warnings.warn(f'sd_animation_optical_flow_b200.{who}: falling back to separate convolution and ReLU kernels: {why}', RuntimeWarning,
                  stacklevel=3)


def _w(conv, pad_out_to: int | None = None):
    w, b = conv.weight.detach(), conv.bias.detach()
    if pad_out_to is not None and w.shape[0] < pad_out_to:
        extra = pad_out_to - w.shape[0]
        w = torch.cat([w, w.new_zeros((extra, *w.shape[1:]))], 0)
        b = torch.cat([b, b.new_zeros(extra)], 0)
    return w.contiguous(memory_format=CL), b.contiguous(), conv.padding


def _fold_bn(conv, bn):
    """Inference BatchNorm folded into the preceding convolution (exact algebra):
    w' = w * g / sqrt(var + eps),  b' = (b - mean) * g / sqrt(var + eps) + beta."""
    w, b = conv.weight.detach(), conv.bias.detach()
    if isinstance(bn, torch.nn.BatchNorm2d):
        k = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
        w = w * k.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach()) * k + bn.bias.detach()
    return w.contiguous(), b.contiguous()


class FastEncoder:
    """Inference forward of raft.Encoder (basic model; RAFT/core/extractor.py:118-192), channels-last end to end:
    cuDNN's tensor-core convolutions take and produce NHWC, so none of the NCHW<->NHWC transposes eager PyTorch
    wraps around each convolution run (0.86 ms of the round-1 5.9 ms step).  Normalisation is collapsed:

      * fnet (InstanceNorm): bias-free convolutions (a per-channel constant cancels in the norm), then
        `instnorm_stats_nhwc` + `instnorm_apply_nhwc` (csrc/raft_glue.cu); the apply kernel of a unit's second
        convolution also does the residual tail `relu(x + y)` (extractor.py:49-58);
      * cnet (BatchNorm): folded into the conv weights, ReLU fused by cuDNN, residual tail = `add_relu`.
    """

    N_NORMS = 15   # stem + 6 units x 2 + 2 downsample projections

    def __init__(self, enc, dtype: torch.dtype = torch.float32):
        """dtype=torch.float16 (InstanceNorm encoders only): activations and filters in fp16, cuDNN tensor-core convolutions
        with fp32 accumulation, normalisation statistics in fp32/fp64.  fp16 keeps the 11 significant bits TF32 keeps of an
        fp32 operand, so the arithmetic is TF32-equivalent at half the activation bytes (every activation is normalised,
        i.e. O(1..100): no range issue); the result is returned in fp32."""
        self.dtype = dtype
        self.kind = enc.norm_fn
        if self.kind not in ('instance', 'batch', 'none'):
            raise ValueError(f'unsupported norm {self.kind}')
        self.stem = self._layer(enc.conv1, enc.norm1)
        # 3 -> 4 input channels (zero filter plane): cuDNN's tensor-core NHWC kernels need C % 4 == 0; with 3 channels the
        # 7x7 stem runs on a CUDA-core engine (266 us per pair in the round-1 launch list)
        w, b, st, pd = self.stem
        self.stem = (torch.cat([w, w.new_zeros((w.shape[0], 1, *w.shape[2:]))], 1).contiguous(memory_format=CL), b, st, pd)
        self.units = []
        for layer in (enc.layer1, enc.layer2, enc.layer3):
            for u in layer:
                ds = self._layer(u.downsample[0], u.downsample[1]) if u.downsample is not None else None
                self.units.append((self._layer(u.conv1, u.norm1), self._layer(u.conv2, u.norm2), ds))
        self.out = (enc.conv2.weight.detach().to(dtype).contiguous(memory_format=CL), enc.conv2.bias.detach().to(dtype), enc.conv2.stride,
                    enc.conv2.padding)
        self._out_bias32 = enc.conv2.bias.detach().float().view(1, -1, 1, 1)
        if dtype != torch.float32:
            cast = lambda l: None if l is None else (l[0].to(dtype).contiguous(memory_format=CL), l[1].to(dtype), l[2], l[3])
            self.stem = cast(self.stem)
            self.units = [tuple(cast(l) for l in u) for u in self.units]
        self._fused_ok = None
        self._max_c = max(l[0].shape[0] for l in [self.stem] + [x for u in self.units for x in u if x is not None])

    def _layer(self, conv, norm):
        w, b = _fold_bn(conv, norm)
        return (w.contiguous(memory_format=CL), b, conv.stride, conv.padding)

    @staticmethod
    def _cl(y):
        return y if y.is_contiguous(memory_format=CL) else y.contiguous(memory_format=CL)

    def _conv_norm(self, x, layer, relu: bool, residual=None):
        """relu?(norm(conv(x))), then relu(residual + .) when a residual is given."""
        w, b, stride, pad = layer
        if self.kind == 'instance':
            y = self._cl(F.conv2d(x, w, None, stride=stride, padding=pad))
            stats = self._stats[self._slot]
            self._slot += 1
            return ops.instnorm_nhwc(y, stats, relu=relu, residual=residual)
        if relu:
            y = None
            if self._fused_ok is not False:
                try:
                    y = torch.cudnn_convolution_relu(x, w, b, tuple(stride), tuple(pad), (1, 1), 1)
                    if self._fused_ok is None:
                        ref = F.relu(F.conv2d(x, w, b, stride=stride, padding=pad))
                        tol = 1e-3 if x.dtype == torch.float32 else 2e-2
                        self._fused_ok = bool(torch.allclose(y, ref, atol=tol, rtol=tol))
                        if not self._fused_ok:
                            _warn_fallback('FastEncoder', 'cudnn_convolution_relu disagrees with conv2d + relu')
                            y = ref
                except RuntimeError as ex:
                    self._fused_ok = False
                    _warn_fallback('FastEncoder', f'cudnn_convolution_relu is unavailable ({ex})')
                    y = None
            if y is None:
                y = F.relu_(F.conv2d(x, w, b, stride=stride, padding=pad))
        else:
            y = F.conv2d(x, w, b, stride=stride, padding=pad)
        y = self._cl(y)
        if residual is not None:
            if y.dtype == torch.float32:
                ops.add_relu_(y, residual)
            else:
                y = torch.relu_(y.add_(residual))
        return y

    def __call__(self, x):
        if x.shape[1] == 3:   # callers on the fast path already hand over 4 channels (ops.normalize_pad_u8(.., channels=4))
            x = F.pad(x, (0, 0, 0, 0, 0, 1))
        x = x.to(self.dtype).contiguous(memory_format=CL)
        if self.kind == 'instance':
            # one zeroed fp64 scratch for the statistics of every norm layer of this pass
            self._stats = torch.zeros((self.N_NORMS, x.shape[0] * self._max_c * 2), dtype=torch.float64, device=x.device)
            self._slot = 0
        x = self._conv_norm(x, self.stem, True)
        for c1, c2, ds in self.units:
            y = self._conv_norm(x, c1, True)
            if ds is not None:
                x = self._conv_norm(x, ds, False)
            x = self._conv_norm(y, c2, True, residual=x)
        w, b, stride, pad = self.out
        if self.dtype == torch.float32:
            return self._cl(F.conv2d(x, w, b, stride=stride, padding=pad))
        # fp16 path: bias-free convolution, then ONE element-wise kernel that adds the fp32 bias and widens to fp32
        # (a biased fp16 conv + .float() were a bias kernel of 11 us and a copy of 12 us)
        y = self._cl(F.conv2d(x, w, None, stride=stride, padding=pad))
        return self._cl(torch.add(y, self._out_bias32))


class KeyFeatures:
    """What a key frame contributes to every pair that uses it as image2 (SURVEY §8a flow-direction note): its feature map
    and, on the 16-bit correlation path, the prepared TARGET operand (the avg-pooled fmap2 levels, built once)."""

    def __init__(self, fmap2_nhwc: torch.Tensor, target):
        self.fmap2, self.target = fmap2_nhwc, target


class FastRaft:
    def __init__(self, model, corr_precision: str = 'fp16', side_streams: bool = True, own_convf1: bool = True,
                 own_fh2: bool = True, corr_storage: str | None = None, tc_gru: bool = False, fnet_fp16: bool = True, cnet_fp16: bool = True, loop_fp16: bool = True,
                 defer_coords: bool = True, convf1_gemm: bool = True):
        """side_streams / own_convf1 / own_fh2 switch the side-stream branches and the two hand-written
        convolutions off (cuDNN + flow_update instead): A/B switches for bench.py, results are identical.
        corr_storage: 'fp16' / 'fp32' pyramid storage (default: fp16 with 16-bit correlation operands, else fp32)."""
        self.side_streams, self.own_convf1, self.own_fh2 = side_streams, own_convf1, own_fh2
        self.corr_storage = corr_storage or ('fp16' if corr_precision in ('fp16', 'bf16') else 'fp32')
        # tc_gru: the SepConvGRU on tcgen05 (csrc/conv_tc.cu: fp16 operands, gate arithmetic in the epilogue) instead of
        # cuDNN TF32 convolutions + element-wise glue kernels.  Correct and parity-tested, but OFF by default: measured on the
        # B200 (tools/gru_bench.py, in-graph, 768x512 batch 1) gru_zr_tc 18.4 us vs cuDNN 11.8 + gru_rh 3.2, gru_q_tc 12.1 vs
        # 6.2 + 3.9, whole step 4.40 vs 3.91 ms (inside the step, warm-cache ncu: 25.6 / 18.7 us vs 13.2 + 5.0 / 9.1 + 5.6);
        # at batch 8 146 vs 82 us.  cuDNN's kernels for these shapes are 2-SM
        # (cta_group::2) tiles with cluster multicast, which halve the L2 -> SM operand traffic this one-CTA-per-tile kernel
        # pays in full (26 B/clk/SM through TMA); profiles/README.md has the numbers and what closing the gap needs.
        self.tc_gru = bool(tc_gru)
        if model.small:
            raise ValueError('FastRaft implements the basic RAFT model (the one the ofgen scripts use)')
        self.model = model
        self.corr_precision = corr_precision
        ub = model.update_block
        e, g, fh = ub.encoder, ub.gru, ub.flow_head
        self.convc1, self.convc2 = _w(e.convc1), _w(e.convc2)
        self.convf1, self.convf2 = _w(e.convf1), _w(e.convf2)
        self.conv = _w(e.conv, pad_out_to=128)           # 126 -> 128 filters (two zero filters) keeps rows 16-byte aligned
        # the same convolution split over its two input groups [cor(192) | flo(64)] (update.py:94-95): each branch
        # convolves its own group on its own stream and the partial results are summed by relu_scatter
        wc = self.conv[0]
        self.conv_cor = (wc[:, :192].contiguous(memory_format=CL), None, e.conv.padding)
        self.conv_flo = (wc[:, 192:].contiguous(memory_format=CL), None, e.conv.padding)
        # GRU input hx = [h | inp | motion | flow] (update.py:47).  `inp` (the context features) does not change over the
        # iterations, so its share of convz/convr/convq is convolved ONCE per pair into per-pixel bias maps (exact by
        # linearity) and the per-iteration convolutions run on [h | motion | flow] only: 256 instead of 384 input channels.
        hd, cd = model.hidden_dim, model.context_dim
        keep = list(range(hd)) + list(range(hd + cd, hd + cd + 128))
        ctx = list(range(hd, hd + cd))
        self.zr, self.q, self.zr_ctx, self.q_ctx = [], [], [], []
        for p in ('1', '2'):
            cz, cr, cq = getattr(g, 'convz' + p), getattr(g, 'convr' + p), getattr(g, 'convq' + p)
            wzr = torch.cat([cz.weight.detach(), cr.weight.detach()], 0)
            bzr = torch.cat([cz.bias.detach(), cr.bias.detach()], 0).contiguous()
            wq, bq = cq.weight.detach(), cq.bias.detach().contiguous()
            # convq's input [r*h | x]: only the first `hd` channels depend on r.  The [motion | flow] share of convq is
            # appended to the z|r convolution as 128 more filters (their h-channel taps are zero), so the convolution
            # that has to wait for r shrinks to hd -> hd channels and needs no [r*h | x] concatenation any more.
            wq_x = wq[:, keep].clone()
            wq_x[:, :hd] = 0
            self.zr.append((torch.cat([wzr[:, keep], wq_x], 0).contiguous(memory_format=CL), bzr, cz.padding))
            self.q.append((wq[:, :hd].contiguous(memory_format=CL), bq, cq.padding))
            self.zr_ctx.append((wzr[:, ctx].contiguous(memory_format=CL), bzr, cz.padding))
            self.q_ctx.append((wq[:, ctx].contiguous(memory_format=CL), bq, cq.padding))
        # K-major fp16 copies of the per-iteration GRU filters for the tcgen05 path
        self.zr16 = [ops.gru_weights16(w[0]) for w in self.zr]
        self.q16 = [ops.gru_weights16(w[0]) for w in self.q]
        self.fh1, self.fh2 = _w(fh.conv1), _w(fh.conv2)
        self.convf1_t = e.convf1.weight.detach().permute(2, 3, 1, 0).contiguous()      # [7,7,2,128] for conv7x7_c2_relu
        self.fh2_t = fh.conv2.weight.detach().permute(2, 3, 0, 1).contiguous()         # [3,3,2,256] for flowhead2_update
        self._fh2_bias = tuple(float(v) for v in fh.conv2.bias.detach().cpu().tolist())
        self.mask0, self.mask2 = _w(ub.mask[0]), _w(ub.mask[2])
        # loop_fp16: the update block's cuDNN convolutions in fp16 (tensor-op, fp32 accumulation), activations between them in
        # fp16 through the glue kernels of csrc/raft_glue16.cu; hidden-state master copy, coordinates, flow and the per-pair
        # bias maps stay fp32.  Same 11-bit operand precision as the TF32 path, 93 instead of 113 us of convolutions per
        # iteration at 768x512.  Only under cuDNN's TF32 default (with TF32 off the caller wants fp32 convolutions).
        self.loop_fp16 = bool(loop_fp16)
        if self.loop_fp16:
            h16 = lambda t: t.to(torch.float16)
            def half_wbp(wbp, pad_in_to=None):
                w, b, pad = wbp
                if pad_in_to is not None and w.shape[1] < pad_in_to:
                    w = torch.cat([w, w.new_zeros((w.shape[0], pad_in_to - w.shape[1], *w.shape[2:]))], 1)
                return (h16(w).contiguous(memory_format=CL), None if b is None else h16(b), pad)
            self.corr_ch16 = 328                                           # 324 lookup channels padded to a multiple of 8
            self.convc1_16 = half_wbp(self.convc1, self.corr_ch16)
            self.convc2_16, self.convf2_16 = half_wbp(self.convc2), half_wbp(self.convf2)
            self.conv_cor_16, self.conv_flo_16 = half_wbp(self.conv_cor), half_wbp(self.conv_flo)
            self.zr_w16 = [half_wbp((w, None, pad)) for (w, _, pad) in self.zr]
            self.q_w16 = [half_wbp((w, None, pad)) for (w, _, pad) in self.q]
            self.fh1_16 = half_wbp(self.fh1)
            self.mask0_16, self.mask2_16 = half_wbp(self.mask0), half_wbp((self.mask2[0], None, self.mask2[2]))
            # convf1 as a 1x1 tensor-core convolution over fp16 im2col rows (ops.flow_im2col7_h): [tap][ci] x (hi | lo) + padding
            self.convf1_k16 = 200
            self.convf1_gemm16 = (ops.im2col7_weight(e.convf1.weight, self.convf1_k16), h16(e.convf1.bias.detach()), (0, 0))
        self.hidden = model.hidden_dim
        self.cdim = model.context_dim
        self._side = {}
        self.defer_coords = bool(defer_coords)
        self.convf1_gemm = bool(convf1_gemm)     # convf1 = im2col kernel + cuDNN 1x1 tensor-core convolution (fp16 loop with deferred coords)
        self._fused_relu_ok = None
        # fnet_fp16: the feature encoder's activations / cuDNN convolutions in fp16 (TF32-equivalent operand precision, half
        # the bytes through the InstanceNorm kernels); its output feeds the correlation, whose operands are fp16 anyway.
        # Only together with cuDNN's TF32 default: with TF32 switched off the caller wants true fp32 convolutions.
        self.fnet_fp16 = bool(fnet_fp16) and model.fnet.norm_fn == 'instance'
        self._fnet32 = FastEncoder(model.fnet)
        self._fnet16 = FastEncoder(model.fnet, torch.float16) if self.fnet_fp16 else None
        # the context encoder the same way (BatchNorm folded into fp16 filters): its output only passes through tanh / relu
        self.cnet_fp16 = bool(cnet_fp16) and model.cnet.norm_fn == 'batch'
        self._cnet32 = FastEncoder(model.cnet)
        self._cnet16 = FastEncoder(model.cnet, torch.float16) if self.cnet_fp16 else None

    def cnet(self, x):
        use16 = self._cnet16 is not None and torch.backends.cudnn.allow_tf32
        return (self._cnet16 if use16 else self._cnet32)(x)

    def fnet(self, x):
        use16 = self._fnet16 is not None and torch.backends.cudnn.allow_tf32
        return (self._fnet16 if use16 else self._fnet32)(x)

    # ---- cuDNN convolutions on dense NHWC buffers ---------------------------------------------------
    @staticmethod
    def _conv(x_nhwc: torch.Tensor, wbp, bias: bool = False) -> torch.Tensor:
        """cuDNN convolution on a dense NHWC buffer.  By default WITHOUT the bias: the consumer glue kernel adds it
        (a separate strided bias pass costs as much as the convolution at this size)."""
        w, b, pad = wbp
        y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w, b if bias else None, padding=pad)
        if not y.is_contiguous(memory_format=CL):
            y = y.contiguous(memory_format=CL)
        return y.permute(0, 2, 3, 1)

    def _conv_relu(self, x_nhwc: torch.Tensor, wbp) -> torch.Tensor:
        w, b, pad = wbp
        x = x_nhwc.permute(0, 3, 1, 2)
        if self._fused_relu_ok is not False:
            try:
                y = torch.cudnn_convolution_relu(x, w, b, (1, 1), tuple(pad), (1, 1), 1)
                if self._fused_relu_ok is None:
                    ref = F.relu(F.conv2d(x, w, b, padding=pad))
                    self._fused_relu_ok = bool(torch.allclose(y, ref, atol=1e-3, rtol=1e-3))
                    if not self._fused_relu_ok:
                        _warn_fallback('FastRaft', 'cudnn_convolution_relu disagrees with conv2d + relu')
                        y = ref
            except RuntimeError as ex:
                self._fused_relu_ok = False
                _warn_fallback('FastRaft', f'cudnn_convolution_relu is unavailable ({ex})')
                y = F.relu_(F.conv2d(x, w, b, padding=pad))
        else:
            y = F.relu_(F.conv2d(x, w, b, padding=pad))
        if not y.is_contiguous(memory_format=CL):
            y = y.contiguous(memory_format=CL)
        return y.permute(0, 2, 3, 1)

    def _conv16_relu(self, x16_nhwc: torch.Tensor, wbp) -> torch.Tensor:
        """fp16 cuDNN convolution + bias + ReLU on a dense NHWC fp16 buffer (fused by cuDNN)."""
        w, b, pad = wbp
        y = torch.cudnn_convolution_relu(x16_nhwc.permute(0, 3, 1, 2), w, b, (1, 1), tuple(pad), (1, 1), 1)
        if not y.is_contiguous(memory_format=CL):
            y = y.contiguous(memory_format=CL)
        return y.permute(0, 2, 3, 1)

    @torch.no_grad()
    def _iterate_fp16(self, iters, B, h, w, dev, main, side, pyr, coords1, flow, H, ZRMAP, QMAP, fh_scratch):
        """The update loop with fp16 activations (see loop_fp16 in __init__): same schedule as the fp32 loop in forward()."""
        hd = self.hidden
        f16 = torch.float16
        HX16 = torch.empty((B, h, w, hd + 128), device=dev, dtype=f16)    # [h | motion | flow], the GRU input
        RH16 = torch.empty((B, h, w, hd), device=dev, dtype=f16)          # r*h
        H16 = torch.empty((B, h, w, hd), device=dev, dtype=f16)           # dense fp16 copy of h for the flow / mask heads
        MF16 = torch.empty((B, h, w, 128), device=dev, dtype=f16)
        F1_16 = torch.empty((B, h, w, 128), device=dev, dtype=f16)
        I2C16 = torch.empty((B, h, w, self.convf1_k16), device=dev, dtype=f16) if (self.defer_coords and self.convf1_gemm) else None
        corr16 = torch.empty((B, h, w, self.corr_ch16), device=dev, dtype=f16)
        H16.copy_(H)
        HX16[..., :hd] = H16
        # Deferred coords update: the flow head leaves its tap products in fh_scratch and the NEXT iteration's consumers of the
        # coordinates (lookup on this stream, convf1 on the side stream) add bias + the 9-neighbour sum themselves -- one kernel
        # and one dependency less per iteration.  The coordinates ping-pong between two buffers because convf1 reads the old
        # ones while the lookup writes the new ones.
        defer = self.defer_coords and iters > 0
        cbuf = [coords1, torch.empty_like(coords1)] if defer else None
        for it in range(iters):
            if defer:
                cin, cout = cbuf[it % 2], cbuf[(it + 1) % 2]
                taps = fh_scratch if it > 0 else None
            side.wait_stream(main)
            with torch.cuda.stream(side):                                 # flow branch (update.py:93-94)
                if defer and self.convf1_gemm:
                    ops.flow_im2col7_h(cin, taps, self._fh2_bias, I2C16)
                    f1 = self._conv16_relu(I2C16, self.convf1_gemm16)
                elif defer:
                    f1 = ops.conv7x7_c2_relu_coords_h(cin, taps, self._fh2_bias, self.convf1_t, self.convf1[1], F1_16)
                else:
                    f1 = ops.conv7x7_c2_relu_h(flow, self.convf1_t, self.convf1[1], F1_16)
                f2 = self._conv16_relu(f1, self.convf2_16)
                del f1
                MF16.copy_(self._conv(f2, self.conv_flo_16))
                del f2
            if defer:                                                     # correlation branch (update.py:91-92)
                ops.corr_lookup_gather_nhwc_h(pyr, cin, taps, self._fh2_bias, cout, flow, corr16)
            else:
                ops.corr_lookup_nhwc_h(pyr, coords1, corr16)
            c2 = self._conv16_relu(self._conv16_relu(corr16, self.convc1_16), self.convc2_16)
            mc = self._conv(c2, self.conv_cor_16)
            main.wait_stream(side)
            ops.motion_tail16_h(mc, MF16, self.conv[1], flow, HX16)
            for p in (0, 1):                                              # SepConvGRU: 1x5 then 5x1 (update.py:45-60)
                zr = self._conv(HX16, self.zr_w16[p])                     # [z | r | x-share of q] fp16
                ops.gru_rh_h(zr, ZRMAP[p], H, RH16)
                q = self._conv(RH16, self.q_w16[p])
                ops.gru_update_h(zr, ZRMAP[p], q, QMAP[p], H, HX16, H16 if p == 1 else None)
            if defer:
                ops.flowhead2_taps_h(self._conv16_relu(H16, self.fh1_16), self.fh2_t, fh_scratch)
            else:
                ops.flowhead2_update_h(self._conv16_relu(H16, self.fh1_16), self.fh2_t, self._fh2_bias, coords1, flow, fh_scratch)
        if defer:                                                         # the last iteration's update
            ops.flowhead2_gather_update(fh_scratch, self._fh2_bias, cbuf[iters % 2], flow)
        mask = self._conv(self._conv16_relu(H16, self.mask0_16), self.mask2_16).float()
        return flow, ops.convex_upsample(mask.contiguous(), flow, 0.25, mask_bias=self.mask2[1])

    def _side_stream(self, dev):
        st = self._side.get(dev)
        if st is None:
            st = self._side[dev] = torch.cuda.Stream(device=dev)
        return st

    @torch.no_grad()
    def encode_key(self, image2: torch.Tensor, normalized: bool = False) -> KeyFeatures:
        """Feature map (+ prepared correlation target) of ONE key frame [1,3|4,H,W]: fnet's InstanceNorm is per image, so
        encoding the key alone gives exactly the features it has inside a pair batch."""
        im2 = image2 if normalized else (2 * (image2 / 255.0) - 1.0).contiguous()
        fmap2 = _to_nhwc(self.fnet(im2))
        target = ops.CorrTarget(fmap2, 4, self.corr_precision) if self.corr_precision in ('fp16', 'bf16') else None
        return KeyFeatures(fmap2, target)

    @torch.no_grad()
    def forward(self, image1: torch.Tensor, image2: torch.Tensor | None, iters: int = 20, normalized: bool = False,
                key: KeyFeatures | None = None, sequence: bool = False):
        """image1/2: [B,3,H,W] float 0..255 (H, W multiples of 8), or with normalized=True already 2*(x/255)-1 (any memory
        format; ops.normalize_pad_u8 hands over channels-last).  Returns (flow_low [B,h,w,2], flow_up [B,H,W,2]).
        key: the encoded key frame serving as image2 of all B pairs (then `image2` is ignored): fnet(key) and the pooled
        correlation operands are not recomputed per pair.
        sequence: image1 holds the B+1 frames of a clip and the pairs are (frame i, frame i+1) -- the loop of ofgen.py, where
        frame i+1 is image2 of one pair and image1 of the next: the feature encoder runs once per FRAME (B+1 images
        instead of 2B), the context encoder on frames 0..B-1.  `image2` is ignored.

        Two independent chains run on a side stream (fork/join with events, so a CUDA-graph capture records them as
        parallel branches): the context encoder next to the feature encoder + correlation pyramid, and in every
        iteration the flow branch of the motion encoder (convf1, convf2) next to lookup -> convc1 -> convc2.  Every
        buffer that crosses the streams is allocated before the fork on the main stream; the side stream's own
        temporaries are allocated and freed in stream order on that stream."""
        B, _, Hh, Ww = image1.shape
        if sequence:
            if key is not None or B < 2:
                raise RuntimeError('sequence mode takes the B+1 >= 2 frames of a clip in image1 and no key')
            B -= 1
        h, w = Hh // 8, Ww // 8
        hd = self.hidden
        dev = image1.device
        cdim = self.cdim                                                  # context ("inp") channels
        xc = 128                                                          # per-iteration x = [motion(126) | flow(2)]; inp is folded into bias maps
        main = torch.cuda.current_stream(dev)
        side = self._side_stream(dev) if self.side_streams else main
        if normalized:
            im1, im2 = image1, image2
        else:
            im1 = (2 * (image1 / 255.0) - 1.0).contiguous()
            im2 = (2 * (image2 / 255.0) - 1.0).contiguous() if key is None and not sequence else None
        frames = im1
        if sequence:
            im1 = frames[:B]
        tc = self.tc_gru
        H = torch.empty((B, h, w, hd), device=dev)                        # hidden state, dense [B,h,w,128] (fp32 master copy)
        if tc:                                                            # tcgen05 GRU: the conv operands exist only in fp16
            HX16 = torch.empty((B, h, w, hd + xc), device=dev, dtype=torch.float16)   # [h | motion | flow]
            RH16 = torch.empty((B, h, w, hd), device=dev, dtype=torch.float16)        # r*h
            Z = torch.empty((B, h, w, hd), device=dev)                    # sigmoid(z)
            QX = torch.empty((B, h, w, hd), device=dev)                   # r-independent share of q
            HX = RH = None
        else:
            HX = torch.empty((B, h, w, hd + xc), device=dev)             # [h | motion | flow]      (update.py:47 minus inp)
            RH = torch.empty((B, h, w, hd), device=dev)                   # r*h                      (update.py:50)
        ZRMAP = [torch.empty((B, h, w, 2 * hd), device=dev) for _ in (0, 1)]   # bias + conv(inp) of convz|convr, per GRU pass
        QMAP = [torch.empty((B, h, w, hd), device=dev) for _ in (0, 1)]        # bias + conv(inp) of convq
        MF = torch.empty((B, h, w, 128), device=dev)                      # flow-branch share of the motion-encoder output conv
        corr = torch.empty((B, h, w, 4 * 81), device=dev)
        flow = torch.empty((B, h, w, 2), device=dev)
        fh_scratch = torch.empty((B * h * w * 18,), device=dev)          # tap products of the flow head's second conv
        mo = hd                                                           # motion-feature slot
        fo = mo + 126                                                     # flow slot
        ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing='ij')
        coords1 = torch.stack([xs, ys], -1).float()[None].repeat(B, 1, 1, 1).contiguous()
        ops.flow_update(None, coords1, flow, HX, fo, None, 0)             # flow = 0 into every slot

        side.wait_stream(main)
        with torch.cuda.stream(side):                                     # ---- context encoder branch
            cn = self.cnet(im1).permute(0, 2, 3, 1)
            torch.tanh(cn[..., :hd], out=H)
            if tc:
                HX16[..., :hd] = H
            else:
                HX[..., :hd] = H
            inp = torch.relu(cn[..., hd:]).contiguous()
            for p in (0, 1):                                              # once per pair: the context share of the GRU convolutions
                ZRMAP[p].copy_(self._conv(inp, self.zr_ctx[p], bias=True))
                QMAP[p].copy_(self._conv(inp, self.q_ctx[p], bias=True))
            del cn, inp
        if sequence:                                                      # ---- feature encoder once per frame
            fmaps = _to_nhwc(self.fnet(frames))
            pyr = ops.corr_volume_pyramid(fmaps[:B], fmaps[1:], 4, self.corr_precision, self.corr_storage)
        elif key is None:                                                 # ---- feature encoder + all-pairs volume
            fmaps = self.fnet(torch.cat([im1, im2], 0))
            pyr = ops.corr_volume_pyramid(_to_nhwc(fmaps[:B]), _to_nhwc(fmaps[B:]), 4, self.corr_precision, self.corr_storage)
        else:                                                             # key frame: its features / operands exist already
            fmaps = self.fnet(im1)
            if key.target is not None:
                pyr = ops.corr_volume_pyramid(_to_nhwc(fmaps), None, 4, self.corr_precision, self.corr_storage, target=key.target)
            else:
                pyr = ops.corr_volume_pyramid(_to_nhwc(fmaps), key.fmap2.expand(B, -1, -1, -1).contiguous(), 4,
                                              self.corr_precision, self.corr_storage)
        del fmaps
        main.wait_stream(side)

        if self.loop_fp16 and not tc and torch.backends.cudnn.allow_tf32:
            return self._iterate_fp16(iters, B, h, w, dev, main, side, pyr, coords1, flow, H, ZRMAP, QMAP, fh_scratch)
        for _ in range(iters):
            side.wait_stream(main)
            with torch.cuda.stream(side):                                 # flow branch (update.py:93-94)
                f1 = ops.conv7x7_c2_relu(flow, self.convf1_t, self.convf1[1]) if self.own_convf1 else self._conv_relu(flow, self.convf1)
                f2 = self._conv_relu(f1, self.convf2)                     # flo (64 channels)
                MF.copy_(self._conv(f2, self.conv_flo))                   # its share of conv([cor | flo]) (update.py:95)
                del f1, f2
            ops.corr_lookup_nhwc(pyr, coords1, 4, corr)                   # correlation branch (update.py:91-92)
            c2 = self._conv_relu(self._conv_relu(corr, self.convc1), self.convc2)   # cor (192 channels)
            mc = self._conv(c2, self.conv_cor)                            # 126 (+2 zero) channels, cor share
            main.wait_stream(side)
            if tc:
                ops.motion_tail16(mc, MF, self.conv[1], flow, HX16)       # relu(conv([cor | flo])) | flow -> fp16 GRU input
                for p in (0, 1):                                          # SepConvGRU: 1x5 then 5x1 (update.py:45-60)
                    ops.gru_zr_tc(HX16, self.zr16[p], ZRMAP[p], H, p == 0, Z, RH16, QX)
                    ops.gru_q_tc(RH16, self.q16[p], QMAP[p], QX, Z, p == 0, H, HX16)
            else:
                ops.relu_scatter(mc, HX, mo, c_valid=126, bias=self.conv[1], src2=MF)
                for p in (0, 1):                                          # SepConvGRU: 1x5 then 5x1 (update.py:45-60)
                    zr = self._conv(HX, self.zr[p])                       # [z | r | x-share of q], 3*hd channels
                    ops.gru_rh(zr, H, RH, bias_zr=ZRMAP[p])
                    q = self._conv(RH, self.q[p])                         # r*h share of q
                    ops.gru_update(zr, q, H, HX, bias_zr=ZRMAP[p], bias_q=QMAP[p])
            if self.own_fh2:
                ops.flowhead2_update(self._conv_relu(H, self.fh1), self.fh2_t, self._fh2_bias, coords1, flow, HX, fo, None, 0,
                                     scratch=fh_scratch)
            else:
                delta = self._conv(self._conv_relu(H, self.fh1), self.fh2)
                ops.flow_update(delta.contiguous(), coords1, flow, HX, fo, None, 0, delta_bias=self._fh2_bias)
        mask = self._conv(self._conv_relu(H, self.mask0), self.mask2)
        return flow, ops.convex_upsample(mask.contiguous(), flow, 0.25, mask_bias=self.mask2[1])
