"""Hand-scheduled channels-last RAFT forward (inference, basic model): the same arithmetic as
`raft.RAFT.forward(test_mode=True)` / the reference's RAFT/core/raft.py:86-144, re-plumbed so that one GRU
iteration is ~25 kernels instead of ~100:

  * every activation of the update block lives in a dense NHWC buffer, so cuDNN's tensor-core convolutions
    run without the NCHW<->NHWC transposes eager PyTorch wraps around each of them;
  * the `torch.cat`s of update.py (hx = [h, x], [r*h, x], [cor, flo], [out, flow]) are persistent
    concatenated buffers (HX, RHX, CF) whose channel slices are written in place by the producers;
  * convz and convr (same input) are one convolution with concatenated filters;
  * the element-wise work between convolutions is four hand-written kernels (csrc/raft_glue.cu), the
    correlation lookup writes channels-last directly (csrc/corr_lookup.cu), and the convex 8x upsample
    is one kernel producing the [B,H,W,2] flow the warp kernel consumes.

The encoders (fnet, cnet) run as the regular PyTorch modules.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .corr import _to_nhwc

CL = torch.channels_last


def _w(conv, pad_out_to: int | None = None):
    w, b = conv.weight.detach(), conv.bias.detach()
    if pad_out_to is not None and w.shape[0] < pad_out_to:
        extra = pad_out_to - w.shape[0]
        w = torch.cat([w, w.new_zeros((extra, *w.shape[1:]))], 0)
        b = torch.cat([b, b.new_zeros(extra)], 0)
    return w.contiguous(memory_format=CL), b.contiguous(), conv.padding


class FastRaft:
    def __init__(self, model, corr_precision: str = 'fp16'):
        if model.small:
            raise ValueError('FastRaft implements the basic RAFT model (the one the ofgen scripts use)')
        self.model = model
        self.corr_precision = corr_precision
        ub = model.update_block
        e, g, fh = ub.encoder, ub.gru, ub.flow_head
        self.convc1, self.convc2 = _w(e.convc1), _w(e.convc2)
        self.convf1, self.convf2 = _w(e.convf1), _w(e.convf2)
        self.conv = _w(e.conv, pad_out_to=128)           # 126 -> 128 filters (two zero filters) keeps rows 16-byte aligned
        self.zr, self.q = [], []
        for p in ('1', '2'):
            cz, cr, cq = getattr(g, 'convz' + p), getattr(g, 'convr' + p), getattr(g, 'convq' + p)
            wzr = torch.cat([cz.weight.detach(), cr.weight.detach()], 0).contiguous(memory_format=CL)
            bzr = torch.cat([cz.bias.detach(), cr.bias.detach()], 0).contiguous()
            self.zr.append((wzr, bzr, cz.padding))
            self.q.append(_w(cq))
        self.fh1, self.fh2 = _w(fh.conv1), _w(fh.conv2)
        self.mask0, self.mask2 = _w(ub.mask[0]), _w(ub.mask[2])
        self.hidden = model.hidden_dim
        self._fused_relu_ok = None

    # ---- cuDNN convolutions on dense NHWC buffers ---------------------------------------------------
    @staticmethod
    def _conv(x_nhwc: torch.Tensor, wbp) -> torch.Tensor:
        w, b, pad = wbp
        y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w, b, padding=pad)
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
                        y = ref
            except RuntimeError:
                self._fused_relu_ok = False
                y = F.relu_(F.conv2d(x, w, b, padding=pad))
        else:
            y = F.relu_(F.conv2d(x, w, b, padding=pad))
        if not y.is_contiguous(memory_format=CL):
            y = y.contiguous(memory_format=CL)
        return y.permute(0, 2, 3, 1)

    @torch.no_grad()
    def forward(self, image1: torch.Tensor, image2: torch.Tensor, iters: int = 20):
        """image1/2: [B,3,H,W] float 0..255 (H, W multiples of 8).  Returns (flow_low [B,h,w,2], flow_up [B,H,W,2])."""
        m = self.model
        im1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        im2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        fmap1, fmap2 = m.fnet([im1, im2])
        pyr = ops.corr_volume_pyramid(_to_nhwc(fmap1), _to_nhwc(fmap2), 4, self.corr_precision)
        cnet = m.cnet(im1)
        B, _, h, w = cnet.shape
        hd = self.hidden
        dev = cnet.device
        cn = cnet.permute(0, 2, 3, 1)
        H = torch.tanh(cn[..., :hd]).contiguous()                       # hidden state, dense [B,h,w,128]
        inp = torch.relu(cn[..., hd:])
        xc = inp.shape[-1] + 128                                          # x = [inp | motion(126) | flow(2)]
        HX = torch.empty((B, h, w, hd + xc), device=dev)                 # [h | x]        (update.py:47)
        RHX = torch.empty_like(HX)                                        # [r*h | x]      (update.py:50)
        HX[..., :hd] = H
        HX[..., hd:hd + inp.shape[-1]] = inp
        RHX[..., hd:hd + inp.shape[-1]] = inp
        mo = hd + inp.shape[-1]                                           # motion-feature slot
        fo = mo + 126                                                     # flow slot
        ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing='ij')
        coords1 = torch.stack([xs, ys], -1).float()[None].repeat(B, 1, 1, 1).contiguous()
        flow = torch.empty((B, h, w, 2), device=dev)
        ops.flow_update(None, coords1, flow, HX, fo, RHX, fo)             # flow = 0 into every slot
        CF = torch.empty((B, h, w, 256), device=dev)                      # [cor(192) | flo(64)]  (update.py:94)
        corr = torch.empty((B, h, w, 4 * 81), device=dev)
        for _ in range(iters):
            ops.corr_lookup_nhwc(pyr, coords1, 4, corr)
            c2 = self._conv(self._conv_relu(corr, self.convc1), self.convc2)
            ops.relu_scatter(c2, CF, 0)
            f2 = self._conv(self._conv_relu(flow, self.convf1), self.convf2)
            ops.relu_scatter(f2, CF, 192)
            mot = self._conv(CF, self.conv)                               # 126 (+2 zero) channels
            ops.relu_scatter(mot, HX, mo, RHX, mo, c_valid=126)
            for p in (0, 1):                                              # SepConvGRU: 1x5 then 5x1 (update.py:45-60)
                zr = self._conv(HX, self.zr[p])
                ops.gru_rh(zr, H, RHX)
                q = self._conv(RHX, self.q[p])
                ops.gru_update(zr, q, H, HX)
            delta = self._conv(self._conv_relu(H, self.fh1), self.fh2)
            ops.flow_update(delta.contiguous(), coords1, flow, HX, fo, RHX, fo)
        mask = self._conv(self._conv_relu(H, self.mask0), self.mask2)
        return flow, ops.convex_upsample(mask.contiguous(), flow, 0.25)
