"""Our K1/C4 path against the reference's OWN CUDA kernel, compiled unmodified into oracle/_ref/
(recipe: oracle/build_ref.py; sources stay under /root/reference, only the built module travels).

`alt_cuda_corr.forward(fmap1, fmap2, coords, r)` (RAFT/alt_cuda_corr/correlation.cpp:23-33,51-54) is the one
native boundary of the reference.  Tolerance: both sides sum 256 fp32 products per tap in different orders
(the reference in 8 slabs of 32 channels with a global read-modify-write, correlation_kernel.cu:43-114),
so 2e-4 abs on |corr| <= ~80 for unit-variance features (the outputs are unnormalised).
"""
import numpy as np
import pytest
import torch

from oracle import build_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ref_mod():
    if build_ref.built_module_path() is None:
        pytest.skip('oracle/_ref/alt_cuda_corr_ref*.so not built (python oracle/build_ref.py in the authoring container)')
    return build_ref.load()


def _inputs(B, H, W, C, lvl, seed, dev):
    g = torch.Generator(device='cpu').manual_seed(seed)
    f1 = torch.randn((B, H, W, C), generator=g).to(dev)
    f2 = torch.randn((B, H >> lvl, W >> lvl, C), generator=g).to(dev)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    grid = torch.stack([xs, ys], -1).float()[None, None]                      # [1,1,H,W,2] (x,y)
    coords = ((grid + 3.0 * torch.randn((B, 1, H, W, 2), generator=g)) / (1 << lvl)).contiguous().to(dev)
    return f1, f2, coords


@pytest.mark.parametrize('shape', [(1, 24, 32, 256), (2, 16, 40, 64), (1, 96, 64, 256)])
@pytest.mark.parametrize('lvl', [0, 1, 3])
def test_alt_corr_forward_matches_reference_kernel(cuda, ref_mod, shape, lvl):
    from sd_animation_optical_flow_b200 import alt_cuda_corr
    B, H, W, C = shape
    f1, f2, coords = _inputs(B, H, W, C, lvl, 7 * lvl + C, cuda)
    ref, = ref_mod.forward(f1, f2, coords, 4)
    torch.cuda.synchronize()
    ours, = alt_cuda_corr.forward(f1, f2, coords, 4)
    assert ours.shape == ref.shape == (B, 1, 81, H, W)
    assert ours.dtype == ref.dtype == torch.float32
    err = float((ours - ref).abs().max())
    assert err <= 2e-4, f'max abs err vs the reference kernel {err} (max|corr| {float(ref.abs().max())})'


def test_alternate_corr_block_matches_reference_kernel_pipeline(cuda, ref_mod):
    """The whole AlternateCorrBlock.__call__ (corr.py:74-91) built on the reference op vs ours."""
    import torch.nn.functional as F
    from sd_animation_optical_flow_b200.corr import AlternateCorrBlock
    B, C, h, w = 1, 256, 32, 48
    g = torch.Generator(device='cpu').manual_seed(3)
    fmap1 = torch.randn((B, C, h, w), generator=g).to(cuda)
    fmap2 = torch.randn((B, C, h, w), generator=g).to(cuda)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    coords = (torch.stack([xs, ys], 0).float()[None] + 2.5 * torch.randn((B, 2, h, w), generator=g)).to(cuda)
    ours = AlternateCorrBlock(fmap1, fmap2, num_levels=4, radius=4)(coords)
    # reference pipeline, restated around the reference op
    pyr = [(fmap1, fmap2)]
    f1, f2 = fmap1, fmap2
    for _ in range(4):
        f1, f2 = F.avg_pool2d(f1, 2, stride=2), F.avg_pool2d(f2, 2, stride=2)
        pyr.append((f1, f2))
    c = coords.permute(0, 2, 3, 1)
    outs = []
    for i in range(4):
        a = pyr[0][0].permute(0, 2, 3, 1).contiguous()
        b = pyr[i][1].permute(0, 2, 3, 1).contiguous()
        ci = (c / 2 ** i).reshape(B, 1, h, w, 2).contiguous()
        o, = ref_mod.forward(a, b, ci, 4)
        outs.append(o.squeeze(1))
    ref = torch.stack(outs, 1).reshape(B, -1, h, w) / torch.sqrt(torch.tensor(float(C)))
    assert ours.shape == ref.shape
    err = float((ours - ref).abs().max())
    assert err <= 2e-5, f'AlternateCorrBlock vs reference-kernel pipeline: {err}'


def test_reference_kernel_agrees_with_oracle(cuda, ref_mod):
    """Pins the NumPy restatement (oracle/corr_oracle.py::alt_corr_forward) to the reference kernel itself."""
    from oracle import corr_oracle as co
    B, H, W, C = 1, 12, 20, 64
    f1, f2, coords = _inputs(B, H, W, C, 0, 11, cuda)
    ref, = ref_mod.forward(f1, f2, coords, 4)
    want = co.alt_corr_forward(f1.cpu().numpy(), f2.cpu().numpy(), coords.cpu().numpy(), 4)
    np.testing.assert_allclose(ref.cpu().numpy(), want, rtol=0, atol=2e-4)
