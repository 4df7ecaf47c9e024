"""Drop-in for the reference's `pdcnet_of.py`: same names, argument meaning and return
types, numpy in / numpy out, with the arithmetic on the B200.

    from sd_animation_optical_flow_b200.pdcnet_of import create_of_algo, warp_frame, warp_frame_latent

* warp_frame          pdcnet_of.py:34-42   cv2.remap(INTER_CUBIC, BORDER_CONSTANT) -> bit-exact CUDA kernel
* warp_frame_latent   pdcnet_of.py:19-32
* PDCNetPlus.calc     pdcnet_of.py:66-75   (+ calc_batch / .to, used at ofgen_keyframe_inpaint.py:555,594)
* create_of_algo      pdcnet_of.py:77-79

PDCNet+ itself (PruneTruong/DenseMatching) is third-party code outside the reference tree; when
that checkout is importable it is used exactly like the reference does, otherwise any object
with `estimate_flow_and_confidence_map(source, target)` can be plugged in (`network=`), e.g.
`engine.RaftFlowConfidence`.  Parity for the PDCNet+ network is unpinned.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('sd_animation_optical_flow_b200 needs a CUDA device (B200); there is no CPU path')
    return torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())


def _h2d(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)


def _d2h(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> fresh host array.  The copy lands in pinned memory (PyTorch's caching host allocator recycles
    the blocks), which avoids the driver's staged pageable copy: 3 MB of flow per 768x512 pair otherwise costs more
    than the warp kernel."""
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out.numpy()


def warp_frame(frame: np.ndarray, flow: np.ndarray, device=None) -> np.ndarray:
    """out[y,x] = cubic(frame, x + flow[y,x,0], y + flow[y,x,1]), constant border 0.
    frame uint8/float32 [H,W] or [H,W,C]; flow float32 [H,W,2]; `flow` is not modified."""
    dev = _device(device)
    fl = _h2d(np.asarray(flow, dtype=np.float32), dev)
    out = ops.warp(_h2d(frame, dev), fl, mode='cv2_cubic', sign=1.0)
    return _d2h(out)


def warp_frame_latent(latent: torch.Tensor, flow: np.ndarray, device=None) -> torch.Tensor:
    """latent [1,C,h,w] (CPU or CUDA) -> CPU tensor [1,C,h,w] (pdcnet_of.py:19-32): cubic-resize to the flow's size, cubic
    warp, cubic-resize back -- all three on the device (csrc/resize.cu restates cv2.resize INTER_CUBIC for float images),
    one H2D of latent + flow and one D2H of the result."""
    dev = _device(device)
    if latent.dim() != 4 or latent.shape[0] != 1:
        raise RuntimeError(f'latent must be [1,C,h,w], got {tuple(latent.shape)}')
    lat = latent.detach().to(dev, torch.float32)[0].permute(1, 2, 0).contiguous()       # 'c h w -> h w c'
    lh, lw = lat.shape[:2]
    h, w = flow.shape[:2]
    fl = _h2d(np.asarray(flow, dtype=np.float32), dev)
    big = ops.resize_cubic(lat, h, w)
    C = big.shape[2]
    if C <= 4:
        warped = ops.warp(big, fl, mode='cv2_cubic', sign=1.0)
    else:                                   # the warp kernels take up to 4 channels per call (cv2.remap's own limit)
        warped = torch.cat([ops.warp(big[:, :, c:c + 4].contiguous(), fl, mode='cv2_cubic', sign=1.0) for c in range(0, C, 4)], 2)
    small = ops.resize_cubic(warped.contiguous(), lh, lw)
    return small.permute(2, 0, 1)[None].contiguous().cpu()


def _load_densematching(ckpt_path: str):
    """Build PDCNet+ exactly as pdcnet_of.py:46-64 does, from a sibling DenseMatching checkout."""
    import sys
    sys.path.append('../DenseMatching')
    from models.PDCNet.PDCNet import PDCNet_vgg16  # noqa: E402  (third-party, not vendored)
    from model_selection import load_network  # noqa: E402
    global_gocor_arguments = {'optim_iter': 6, 'steplength_reg': 0.1, 'train_label_map': False,
                              'apply_query_loss': True, 'reg_kernel_size': 3, 'reg_inter_dim': 16, 'reg_output_dim': 16}
    local_gocor_arguments = {'optim_iter': 14, 'steplength_reg': 0.1}
    network = PDCNet_vgg16(global_corr_type='GlobalGOCor', global_gocor_arguments=global_gocor_arguments,
                           normalize='leakyrelu', same_local_corr_at_all_levels=True,
                           local_corr_type='LocalGOCor', local_gocor_arguments=local_gocor_arguments,
                           local_decoder_type='OpticalFlowEstimatorResidualConnection',
                           global_decoder_type='CMDTopResidualConnection',
                           corr_for_corr_uncertainty_decoder='corr',
                           give_layer_before_flow_to_uncertainty_decoder=True,
                           var_2_plus=520 ** 2, var_2_plus_256=256 ** 2, var_1_minus_plus=1.0, var_2_minus=2.0,
                           make_two_feature_copies=True)
    network = load_network(network, checkpoint_path=ckpt_path).cuda()
    network.eval()
    return network


class PDCNetPlus:
    """Flow + confidence back-end with the reference's protocol (pdcnet_of.py:45-75)."""

    def __init__(self, ckpt_path: str = 'pre_trained_models/PDCNet_plus_m.pth.tar', network=None, device=None) -> None:
        self.device = _device(device)
        if network is None:
            try:
                network = _load_densematching(ckpt_path)
            except ImportError as e:
                raise ImportError(
                    'PDCNet+ lives in the third-party DenseMatching checkout the reference imports from '
                    "'../DenseMatching' (pdcnet_of.py:6-13); it is not part of the reference tree. Pass "
                    '`network=` (any object with estimate_flow_and_confidence_map(source, target)), e.g. '
                    'sd_animation_optical_flow_b200.engine.RaftFlowConfidence.') from e
        self.network = network

    def to(self, device):
        """The reference calls `.to(device)` on the algo object (ofgen_keyframe_inpaint.py:555)."""
        self.device = _device(device)
        if hasattr(self.network, 'to'):
            self.network = self.network.to(self.device)
        return self

    @torch.no_grad()
    def _estimate(self, src_u8_bhwc: torch.Tensor, tgt_u8_bhwc: torch.Tensor):
        """RGB uint8 [B,H,W,3] on the device -> (flow [B,H,W,2], weight_map [B,K,H,W])."""
        src = src_u8_bhwc.permute(0, 3, 1, 2)
        tgt = tgt_u8_bhwc.permute(0, 3, 1, 2)
        flow, unc = self.network.estimate_flow_and_confidence_map(src, tgt)
        flow = flow.permute(0, 2, 3, 1).float().contiguous()
        return flow, unc['weight_map'].float().contiguous()

    @torch.no_grad()
    def calc(self, frame1: np.ndarray, frame2: np.ndarray):
        """BGR uint8 [H,W,3] x2 -> (flow f32 [H,W,2], confidence f32 [H,W], log_confidence f32 [H,W]),
        fresh host arrays the caller may mutate (pdcnet_of.py:66-75)."""
        dev = self.device
        f1 = _h2d(frame1, dev)[None].flip(-1)  # BGR -> RGB on the device
        f2 = _h2d(frame2, dev)[None].flip(-1)
        flow, wm = self._estimate(f1.contiguous(), f2.contiguous())
        conf, logc = ops.confidence_softmax(wm.to(dev))
        return _d2h(flow[0]), _d2h(conf[0]), _d2h(logc[0])

    @torch.no_grad()
    def calc_batch(self, src: torch.Tensor, tgt: torch.Tensor):
        """RGB uint8 [B,H,W,3] tensors already on the device -> (flow [B,H,W,2], confidence [B,H,W])
        as host arrays, so `ret[si,ti,:,:,0:2] = flow[i]` works as written at
        ofgen_keyframe_inpaint.py:594-599 (the method is called there but missing from the reference)."""
        flow, conf = self.calc_batch_device(src, tgt)
        return _d2h(flow), _d2h(conf)

    @torch.no_grad()
    def calc_batch_device(self, src: torch.Tensor, tgt: torch.Tensor):
        flow, wm = self._estimate(src.to(self.device).contiguous(), tgt.to(self.device).contiguous())
        conf, _ = ops.confidence_softmax(wm)
        return flow, conf


def create_of_algo(ckpt: str = 'pre_trained_models/PDCNet_plus_m.pth.tar', network=None):
    """pdcnet_of.py:77-79."""
    return PDCNetPlus(ckpt, network=network)
