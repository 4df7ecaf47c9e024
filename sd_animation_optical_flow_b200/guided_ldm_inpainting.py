"""Host mirror of the mask/composite preparation inside the reference's `GuidedLDM.img2img_inpaint`
(guided_ldm_inpainting.py:290-309) -- the step that consumes the hot path's outputs (warped frame + inpainting mask)
right before Stable Diffusion.  Only that preparation is here; the diffusion model itself is out of scope.

    image, image_mask, nmask = prepare_inpaint_inputs(image, mask, mask_blur, reference_img)

takes what the reference takes (PIL images or uint8 arrays) and returns what the reference feeds on:
  image      float32 CUDA [1,3,H,W] in [-1,1]   (np.array(image) / 127.5 - 1.0, :299-301)
  image_mask uint8 numpy [H,W]                   (the blurred mask, used for the inpainting condition, :314)
  nmask      float32 CUDA [4,H/8,W/8] in {0,1}   (:304-309)
computed on the device by csrc/blur.cu, bit-exact to Pillow.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .pdcnet_of import _d2h, _device, _h2d


def _as_u8(img, name: str) -> np.ndarray:
    a = np.asarray(img)      # PIL images convert through the array interface
    if a.dtype != np.uint8:
        raise RuntimeError(f'{name} must be an 8-bit image, got dtype {a.dtype}')
    return a


def prepare_inpaint_inputs(image, mask, mask_blur: float = 4, reference_img=None, latent_hw=None, device=None):
    """guided_ldm_inpainting.py:290-309 with `reference_img` given (the only live call site passes one,
    ofgen_pixel_inpaint.py:236-240).  `mask` may be 'L' or RGB (`.convert('L')` of an RGB mask with equal channels is
    the channel itself); latent_hw defaults to (H//8, W//8), the shape of the VAE latent."""
    if reference_img is None:
        raise NotImplementedError('fill_mask_input (no reference image) is a host-side OpenCV inpaint in the reference and out of scope')
    dev = _device(device)
    img = _as_u8(image, 'image')
    ref = _as_u8(reference_img, 'reference_img')
    m = _as_u8(mask, 'mask')
    if m.ndim == 3:
        if not (np.array_equal(m[..., 0], m[..., 1]) and np.array_equal(m[..., 0], m[..., 2])):
            raise RuntimeError('RGB masks must be grey (equal channels): PIL\'s luma transform of a coloured mask is not mirrored')
        m = m[..., 0]
    H, W = m.shape
    if img.shape != (H, W, 3) or ref.shape != (H, W, 3):
        raise RuntimeError(f'image / reference_img must be [{H},{W},3], got {img.shape} / {ref.shape}')
    lh, lw = latent_hw if latent_hw is not None else (H // 8, W // 8)
    out, blurred = ops.mask_blur_composite(_h2d(m, dev)[None], _h2d(img, dev)[None], _h2d(ref, dev)[None], float(mask_blur))
    _, lat = ops.resize_bicubic_u8(blurred, lh, lw, want_latmask=True)
    # tensor / tensor is IEEE division (tensor / python-scalar multiplies by the reciprocal): same bits as numpy's / 127.5
    image_t = (out[0].permute(2, 0, 1).float().div(torch.full((), 127.5, device=dev)) - 1.0)[None]
    return image_t, _d2h(blurred[0]), lat[0]
