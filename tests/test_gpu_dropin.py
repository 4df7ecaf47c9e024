"""The reference-named numpy entry points end to end on the GPU (PDCNetPlus protocol, PDCNetAux pair
cache, confidence_to_mask, the config-3 clip path)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import mask_oracle as mo
from oracle import warp_oracle as wo
from tests import golden_inputs as gi

pytestmark = pytest.mark.gpu


class _FakeNet:
    """Stands in for DenseMatching's PDCNet_vgg16 at the boundary the reference uses
    (estimate_flow_and_confidence_map, pdcnet_of.py:70): deterministic flow + logits."""

    def to(self, device):
        return self

    def estimate_flow_and_confidence_map(self, source, target):
        B, _, H, W = source.shape
        dev = source.device
        ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing='ij')
        flow = torch.stack([2.5 + 0.01 * ys, -1.25 + 0.02 * xs], 0).float()[None].repeat(B, 1, 1, 1)
        flow = flow + 0.001 * (source.float().mean(1, keepdim=True) - target.float().mean(1, keepdim=True))
        wm = torch.stack([0.05 * (xs - W / 2).float(), 0.03 * (ys - H / 2).float()], 0)[None].repeat(B, 1, 1, 1)
        return flow, {'weight_map': wm}


def test_pdcnetplus_calc_protocol(cuda):
    from sd_animation_optical_flow_b200 import ofgen, pdcnet_of
    algo = pdcnet_of.create_of_algo('unused.pth.tar', network=_FakeNet())
    f1, f2 = gi.shifted_pair(64, 80, 3)
    flow, conf, logc = algo.calc(f1, f2)
    assert flow.shape == (64, 80, 2) and conf.shape == (64, 80) and logc.shape == (64, 80)
    assert flow.dtype == conf.dtype == logc.dtype == np.float32
    assert flow.flags.writeable and conf.flags.writeable            # callers mutate them in place
    np.testing.assert_allclose(np.exp(logc), conf, rtol=1e-5)
    # of_calc flavour of ofgen_pixel_inpaint.py:105-118
    fl, c, v, lc = ofgen.of_calc_pdcnet(f1, f2, algo)
    assert np.array_equal(v, mo.travel_distance(fl, c, 0.9))
    # calc_batch as called at ofgen_keyframe_inpaint.py:594-599
    src = torch.from_numpy(np.stack([f1, f2])).to(cuda)
    fb, cb = algo.calc_batch(src, src.flip(0))
    ret = np.zeros((2, 64, 80, 3), np.float32)
    ret[0, :, :, 0:2] = fb[0]
    ret[0, :, :, 2] = cb[0]
    assert fb.shape == (2, 64, 80, 2) and cb.shape == (2, 64, 80)
    with pytest.raises(ImportError, match='DenseMatching'):
        pdcnet_of.PDCNetPlus('missing.pth.tar')


def test_pdcnet_aux_pair_cache(cuda, tmp_path):
    from sd_animation_optical_flow_b200 import ofgen, pdcnet_of
    frames = [gi.texture(48, 64, s) for s in range(5)]
    video = SimpleNamespace(size_hw=(48, 64), get_raw_frame=lambda i: frames[i])
    aux = ofgen.PDCNetAux(pdcnet_of.PDCNetPlus(network=_FakeNet()), str(tmp_path), batch_size=3)
    mat = aux.calculate_pairwise(video, [0, 2, 4])
    assert mat.shape == (3, 3, 48, 64, 3)
    assert (mat[1, 1, :, :, :2] == 0).all() and (mat[1, 1, :, :, 2] == 1).all()      # identity pair
    files = sorted(os.listdir(os.path.join(str(tmp_path), 'pdcnet')))
    assert files[0] == '00000-00002.npy' and len(files) == 6
    assert np.array_equal(np.load(os.path.join(str(tmp_path), 'pdcnet', '00000-00002.npy')), mat[0, 1])
    aux2 = ofgen.PDCNetAux(pdcnet_of.PDCNetPlus(network=_FakeNet()), str(tmp_path))
    assert len(aux2.cached_pair) == 6                                               # rebuilt from disk
    m2o = aux2.calculate_multiple_to_one(video, [0, 2, 4, 1], 2)
    assert m2o.shape == (4, 1, 48, 64, 3) and np.array_equal(m2o[0, 0], mat[0, 1])
    assert (m2o[1, 0, :, :, 2] == 1).all()
    assert np.array_equal(aux2.calcualte_single(video, 4, 2), mat[2, 1])
    assert ofgen.keyframe_conv_pick(mat) == int(np.argmax(mo.keyframe_scores(mat)))
    aux2.purge()
    assert not os.listdir(os.path.join(str(tmp_path), 'pdcnet'))


def test_confidence_to_mask(cuda):
    from sd_animation_optical_flow_b200 import ofgen
    conf, _, _, _, _ = gi.mask_inputs()
    rs = np.random.RandomState(6)
    H, W = conf.shape
    flow = (2 * rs.standard_normal((H, W, 2))).astype(np.float32)
    dist = rs.uniform(0, 3, (H, W)).astype(np.float32)
    ptd = rs.uniform(0, 10, (H, W)).astype(np.float32)
    aux = SimpleNamespace(pixel_travel_dist=ptd.copy(), thres=9.0)
    m = ofgen.confidence_to_mask(conf, flow, dist, aux)
    m_ref, ptd_ref = mo.confidence_to_mask(conf, flow, dist, ptd, 9.0)
    np.testing.assert_allclose(aux.pixel_travel_dist, ptd_ref, atol=1e-5)
    near = np.abs(mo.confidence_to_mask(conf, flow, dist, ptd, 9.0 + 1e-4)[0].astype(int) - m_ref.astype(int)).sum()
    assert (m != m_ref).sum() <= near + 0


def test_raft_flow_confidence_adapter_and_clip_path(cuda):
    """Config 3 in miniature: key frame -> N frames; flow + (forward-backward) confidence on the target
    grid, fused warp + mask + composite, all on the device."""
    from sd_animation_optical_flow_b200 import ops, pdcnet_of
    from sd_animation_optical_flow_b200.engine import RaftEngine, RaftFlowConfidence
    eng = RaftEngine(iters=4, device=cuda)
    algo = pdcnet_of.PDCNetPlus(network=RaftFlowConfidence(eng))
    key, _ = gi.shifted_pair(128, 160, 31)
    frames = np.stack([gi.shifted_pair(128, 160, 31, dx=d, dy=-d)[1] for d in (1, 2, 3)])
    src = torch.from_numpy(np.repeat(key[None], 3, 0)).to(cuda)
    tgt = torch.from_numpy(frames).to(cuda)
    flow, conf = algo.calc_batch_device(src, tgt)
    assert flow.shape == (3, 128, 160, 2) and conf.shape == (3, 128, 160)
    assert float(conf.min()) >= 0 and float(conf.max()) <= 1
    stylised = torch.from_numpy(gi.texture(128, 160, 77)).to(cuda)[None]
    wm = torch.stack([torch.logit(conf.clamp(1e-4, 1 - 1e-4)), torch.zeros_like(conf)], 1).contiguous()
    out, mask = ops.warp_mask_composite(stylised, tgt, flow, wm, 0.95, 7)
    warped = ops.warp(stylised, flow)
    keep = mask <= 127
    assert torch.equal(out[keep], warped[keep]) and torch.equal(out[~keep], tgt[~keep])


def test_pdcnet_aux_device_resident_multiple_to_one(cuda, tmp_path):
    """PDCNetAux.calculate_multiple_to_one_device == the numpy / `.npy`-cache path of calculate_multiple_to_one on the same
    frames (ofgen_keyframe_inpaint.py:602-625), without leaving the device."""
    from sd_animation_optical_flow_b200 import ofgen, pdcnet_of
    from sd_animation_optical_flow_b200.engine import RaftEngine, RaftFlowConfidence
    from tests import golden_inputs as gi
    eng = RaftEngine(checkpoint=None, iters=3, seed=0, flow_head_scale=0.02, device=cuda, use_cuda_graph=False)
    algo = pdcnet_of.PDCNetPlus(network=RaftFlowConfidence(eng))
    aux = ofgen.PDCNetAux(algo, str(tmp_path), batch_size=2, device=cuda)
    canvas = gi.texture(96 + 16, 128 + 16, 5)
    frames = [np.ascontiguousarray(canvas[2 * i:2 * i + 96, 3 * i:3 * i + 128]) for i in range(4)]     # RGB

    class Video:
        size_hw = (96, 128)

        def get_raw_frame(self, i):
            return frames[i][:, :, ::-1]                                                               # BGR like cv2.imread

    ref = aux.calculate_multiple_to_one(Video(), [0, 1, 3], 3)                                        # [3,1,H,W,3] numpy
    dev_frames = torch.from_numpy(np.stack([frames[0], frames[1], frames[3]])).to(cuda)
    got = aux.calculate_multiple_to_one_device(dev_frames, torch.from_numpy(frames[3]).to(cuda), identity=[False, False, True])
    assert got.is_cuda and tuple(got.shape) == ref.shape
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=2e-4)
