"""CPU oracle for the step AFTER the hot path: the mask blur + composite + latent mask that
`GuidedLDM.img2img_inpaint` applies to the warped frame and the inpainting mask before Stable Diffusion runs
(guided_ldm_inpainting.py:290-309):

    image_mask  = mask.convert('L').filter(ImageFilter.GaussianBlur(mask_blur))          (:292-293)
    image       = Image.composite(reference_img, image, image_mask)                        (:298)
    latmask     = around(image_mask.convert('RGB').resize((w/8, h/8)) / 255)[0], tiled x4  (:304-308)

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The arithmetic lives in a third-party dependency that is not in
the reference tree: Pillow (unpinned by the reference; 12.2.0 in this image).  Its published algorithms are restated
here in NumPy and PINNED against Pillow itself by tests/test_oracle_blur.py:

  * GaussianBlur = 3 passes of an "extended" box blur per axis (libImaging/BoxBlur.c): box radius l + a from
    Gwosdek et al., integer box sum * ww + the two outer pixels * fw in 8.24 fixed point, rounded to uint8 after every
    pass, edges replicated;
  * Image.composite / paste with an 'L' mask (libImaging/Paste.c): out = DIV255(dst*(255-m) + src*m) with
    DIV255(a) = ((a+128) >> 8) + (a+128) >> 8;
  * Image.resize default BICUBIC (libImaging/Resample.c): separable, support scaled by the reduction factor
    (antialias), coefficients normalised and quantised to 22 fractional bits, horizontal pass then vertical pass,
    each rounded and clipped to uint8.
"""
from __future__ import annotations

import math

import numpy as np


# ----------------------------------------------------------------------------- GaussianBlur
def gaussian_box_radius(radius: float, passes: int = 3) -> float:
    """BoxBlur.c::_gaussian_blur_radius — float variables, double expressions."""
    f = np.float32
    sigma2 = f(f(radius) * f(radius) / f(passes))
    L = f(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = f(math.floor((float(L) - 1.0) / 2.0))
    a = f(f(2 * l + 1) * f(l * f(l + 1) - f(3 * sigma2)))
    a = f(a / f(6 * f(sigma2 - f(f(l + 1) * f(l + 1)))))
    return float(f(l + a))


def box_weights(float_radius: float):
    """(radius, ww, fw) of ImagingHorizontalBoxBlur: 8.24 fixed-point weights of the inner box and the two outer pixels."""
    fr = np.float32(float_radius)
    radius = int(fr)
    ww = int(np.float32(16777216.0) / np.float32(fr * np.float32(2) + np.float32(1)))
    fw = ((1 << 24) - (radius * 2 + 1) * ww) // 2
    return radius, ww, fw


def box_blur_rows(img: np.ndarray, float_radius: float) -> np.ndarray:
    """One horizontal extended-box pass over every row of a uint8 [H,W] image (edges replicated)."""
    radius, ww, fw = box_weights(float_radius)
    H, W = img.shape
    x = np.arange(W)
    src = img.astype(np.int64)
    acc = np.zeros((H, W), np.int64)
    for d in range(-radius, radius + 1):
        acc += src[:, np.clip(x + d, 0, W - 1)]
    far = src[:, np.clip(x - radius - 1, 0, W - 1)] + src[:, np.clip(x + radius + 1, 0, W - 1)]
    bulk = (acc * ww + far * fw) & 0xffffffff          # UINT32 arithmetic
    return (((bulk + (1 << 23)) & 0xffffffff) >> 24).astype(np.uint8)


def gaussian_blur_u8(img: np.ndarray, radius: float, passes: int = 3) -> np.ndarray:
    """ImageFilter.GaussianBlur(radius) on an 'L' image: `passes` horizontal passes, then `passes` vertical ones."""
    if img.dtype != np.uint8 or img.ndim != 2:
        raise ValueError('uint8 [H,W] image expected')
    fr = gaussian_box_radius(radius, passes)
    out = img
    if fr != 0:
        for _ in range(passes):
            out = box_blur_rows(out, fr)
        out = out.T
        for _ in range(passes):
            out = box_blur_rows(np.ascontiguousarray(out), fr)
        out = np.ascontiguousarray(out.T)
    return out.copy()


# ----------------------------------------------------------------------------- Image.composite
def div255(a: np.ndarray) -> np.ndarray:
    t = a + 128
    return ((t >> 8) + t) >> 8


def composite(image1: np.ndarray, image2: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Image.composite(image1, image2, mask): image2 where mask = 0, image1 where mask = 255, blended between.
    image1/2 uint8 [H,W,C], mask uint8 [H,W]."""
    m = mask.astype(np.int64)[:, :, None]
    return div255(image2.astype(np.int64) * (255 - m) + image1.astype(np.int64) * m).astype(np.uint8)


# ----------------------------------------------------------------------------- Image.resize (BICUBIC)
def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int, support: float = 2.0):
    """Resample.c::precompute_coeffs + normalize_coeffs_8bpc: per output pixel (xmin, integer coefficients)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    sup = support * filterscale
    bounds, coeffs = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - sup + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + sup + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(k)
        k = [v / ww if ww != 0.0 else v for v in k]
        q = [int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22)) for v in k]
        bounds.append(xmin)
        coeffs.append(np.array(q, np.int64))
    return bounds, coeffs


def _resample_axis_last(img: np.ndarray, out_size: int) -> np.ndarray:
    bounds, coeffs = resample_coeffs(img.shape[-1], out_size)
    out = np.empty(img.shape[:-1] + (out_size,), np.uint8)
    src = img.astype(np.int64)
    for xx, (x0, k) in enumerate(zip(bounds, coeffs)):
        ss = (src[..., x0:x0 + len(k)] * k).sum(-1) + (1 << 21)
        out[..., xx] = np.clip(ss >> 22, 0, 255)
    return out


def resize_bicubic_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Image.resize((out_w, out_h)) with the default BICUBIC filter on a uint8 [H,W] (or [H,W,C]) image:
    horizontal pass first, then vertical (Resample.c::ImagingResampleInner)."""
    a = img if img.ndim == 3 else img[:, :, None]
    a = np.moveaxis(a, 2, 0)                                  # [C,H,W]
    if out_w != a.shape[2]:
        a = _resample_axis_last(a, out_w)
    if out_h != a.shape[1]:
        a = np.swapaxes(_resample_axis_last(np.swapaxes(a, 1, 2), out_h), 1, 2)
    a = np.moveaxis(a, 0, 2)
    return np.ascontiguousarray(a if img.ndim == 3 else a[:, :, 0])


# ----------------------------------------------------------------------------- the reference step
def blur_composite_latmask(image: np.ndarray, reference: np.ndarray, mask: np.ndarray, mask_blur: float, latent_hw=None):
    """guided_ldm_inpainting.py:290-308 with reference_img given: (composited uint8 [H,W,3], image_mask uint8 [H,W],
    latmask float32 [4,h,w] in {0,1})."""
    H, W = mask.shape
    image_mask = gaussian_blur_u8(mask, mask_blur)
    out = composite(reference, image, image_mask)
    lh, lw = latent_hw if latent_hw is not None else (H // 8, W // 8)
    small = resize_bicubic_u8(image_mask, lw, lh)
    latmask = np.around(small.astype(np.float32) / 255)
    return out, image_mask, np.tile(latmask[None], (4, 1, 1)).astype(np.float32)
