"""Edge cases through the host mirror + C ABI: empty batches, 1x1 / one-row images, error behaviour
(the reference's CHECK_INPUT, correlation.cpp:19-21: RuntimeError for non-CUDA / non-contiguous tensors)."""
import numpy as np
import pytest
import torch

from oracle import blur_oracle as bo
from oracle import warp_oracle as wo

pytestmark = pytest.mark.gpu


def test_empty_batches_are_no_ops(cuda):
    from sd_animation_optical_flow_b200 import ops
    u8 = lambda *s: torch.empty(s, dtype=torch.uint8, device=cuda)
    f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=cuda)
    assert ops.warp(u8(0, 8, 8, 3), f32(0, 8, 8, 2)).shape == (0, 8, 8, 3)
    assert ops.warp(u8(1, 8, 8, 3), f32(0, 8, 8, 2), 'bilinear').shape == (0, 8, 8, 3)
    out, mask = ops.warp_mask_composite(u8(1, 8, 8, 3), u8(0, 8, 8, 3), f32(0, 8, 8, 2), f32(0, 2, 8, 8), 0.5, 7)
    assert out.shape == (0, 8, 8, 3) and mask.shape == (0, 8, 8)
    out, blurred = ops.mask_blur_composite(u8(0, 8, 8), u8(0, 8, 8, 3), u8(0, 8, 8, 3), 2.0)
    assert out.shape == (0, 8, 8, 3) and blurred.shape == (0, 8, 8)


def test_tiny_images(cuda):
    """1x1 and single-row / single-column frames: every tap of the cubic warp is border or the pixel itself."""
    from sd_animation_optical_flow_b200 import ops
    rs = np.random.RandomState(0)
    for (H, W) in ((1, 1), (1, 37), (41, 1), (2, 3)):
        img = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        flow = (1.5 * rs.standard_normal((H, W, 2))).astype(np.float32)
        out = ops.warp(torch.from_numpy(img).to(cuda), torch.from_numpy(flow).to(cuda)).cpu().numpy()
        assert np.array_equal(out, wo.warp_frame_pdcnet(img, flow)), (H, W)
        m = rs.randint(0, 256, (1, H, W)).astype(np.uint8)
        _, blurred = ops.mask_blur_composite(torch.from_numpy(m).to(cuda), None, None, 2.0)
        assert np.array_equal(blurred[0].cpu().numpy(), bo.gaussian_blur_u8(m[0], 2.0)), (H, W)


def test_error_behaviour_follows_the_reference_op(cuda):
    from sd_animation_optical_flow_b200 import alt_cuda_corr, ops
    img = torch.zeros((8, 8, 3), dtype=torch.uint8, device=cuda)
    flow = torch.zeros((8, 8, 2), device=cuda)
    with pytest.raises(RuntimeError):
        ops.warp(img.cpu(), flow)                                  # not a CUDA tensor
    with pytest.raises(RuntimeError):
        ops.warp(img, flow.double())                               # wrong dtype
    with pytest.raises(RuntimeError):
        ops.warp(img, torch.zeros((8, 8, 4), device=cuda)[..., :2])  # non-contiguous
    with pytest.raises(ValueError):
        ops.warp(img, flow, mode='nearest')
    f = torch.zeros((1, 4, 4, 32), device=cuda)
    c = torch.zeros((1, 1, 4, 4, 2), device=cuda)
    with pytest.raises(RuntimeError):
        alt_cuda_corr.forward(f.cpu(), f, c, 4)
    with pytest.raises(RuntimeError):
        alt_cuda_corr.forward(f.permute(0, 2, 1, 3), f, c, 4)      # non-contiguous (correlation.cpp:20)
    with pytest.raises(RuntimeError):
        ops.mask_blur_composite(torch.zeros((1, 8, 8), dtype=torch.uint8, device=cuda), None, None, -1.0)
