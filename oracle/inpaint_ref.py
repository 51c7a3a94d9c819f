"""ORACLE (test infrastructure, NOT product code) — numpy restatement of the integer image work A1111's img2img does
around the UNet for the request the reference sends (signerf/diffuser/diffuser.py:132-169: mask_blur 4,
inpainting_fill 1, inpaint_full_res 0), SURVEY §8(f) row 1 / Appendix C items 1 and 4:

    mask -> cv2.GaussianBlur (21,1) then (1,21), sigma 4, uint8      modules/processing.py StableDiffusionProcessingImg2Img.init
    mask_for_overlay = clip(2 * blurred, 0, 255)
    latent mask      = round(PIL bicubic resize of the blurred mask to (W/8, H/8) / 255)
    result           = PIL alpha_composite(generated, original with alpha 255 - mask_for_overlay)   (apply_overlay)

PINNED: A1111 (@5ef669de, external) is not importable, but the arithmetic lives in its dependencies, which ARE here:
cv2 4.13 and Pillow 12.2.  tests/golden/make_inpaint_golden.py runs cv2.GaussianBlur / PIL.Image.resize /
Image.paste / alpha_composite exactly as processing.py calls them and stores the outputs in tests/golden/inpaint.npz;
tests/test_oracle_golden.py checks every function below bit for bit against those fixtures.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


# ---------------------------------------------------------------------------------------------- cv2.GaussianBlur (CV_8U)
def gaussian_kernel_q8(ksize: int, sigma: float) -> np.ndarray:
    """OpenCV's bit-exact 8-bit Gaussian kernel (imgproc smooth.dispatch.cpp: getGaussianKernelBitExact +
    getGaussianKernelFixedPoint_ED): exp(-x^2 / 2 sigma^2) normalised in double, then Q8 fixed point with error
    diffusion from the borders inwards; the centre tap takes whatever keeps the sum at exactly 256."""
    assert ksize % 2 == 1 and ksize >= 1
    if ksize == 1:
        return np.array([256], dtype=np.int64)
    r = ksize // 2
    scale2x = -0.5 / (sigma * sigma)
    k = np.array([math.exp(scale2x * (i - r) * (i - r)) for i in range(ksize)], dtype=np.float64)
    k = k / k.sum()
    out = np.zeros(ksize, dtype=np.int64)
    err = 0.0
    total = 0
    for i in range(r):
        adj = k[i] * 256.0 + err
        v = int(np.rint(adj))          # cvRound: round half to even
        err = adj - v
        out[i] = out[ksize - 1 - i] = v
        total += 2 * v
    out[r] = 256 - total
    return out


def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.abs(i) % p
    return np.where(i >= n, p - i, i)


def gaussian_blur_u8_1d(img: np.ndarray, ksize: int, sigma: float, axis: int) -> np.ndarray:
    """cv2.GaussianBlur(img, (ksize, 1) if axis == 1 else (1, ksize), sigma) on uint8, BORDER_REFLECT_101:
    Q8 taps against the uint8 source give a Q8.8 row value, the unit kernel of the other axis keeps it, and the final
    cast rounds half up: (sum + 128) >> 8."""
    assert img.dtype == np.uint8 and img.ndim == 2
    k = gaussian_kernel_q8(ksize, sigma)
    r = ksize // 2
    n = img.shape[axis]
    src = img.astype(np.int64)
    acc = np.zeros(img.shape, dtype=np.int64)
    idx = np.arange(n)
    for t in range(ksize):
        j = _reflect101(idx + t - r, n)
        acc += k[t] * np.take(src, j, axis=axis)
    return ((acc + 128) >> 8).astype(np.uint8)


def a1111_mask_blur(mask_u8: np.ndarray, mask_blur: int = 4) -> Tuple[np.ndarray, np.ndarray]:
    """-> (blurred mask, mask_for_overlay).  processing.py: kernel_size = 2*int(2.5*blur + 0.5) + 1, x pass then y pass,
    each through uint8; np.clip(blurred.astype(float32) * 2, 0, 255).astype(uint8)."""
    m = mask_u8
    if mask_blur > 0:
        ks = 2 * int(2.5 * mask_blur + 0.5) + 1
        m = gaussian_blur_u8_1d(m, ks, float(mask_blur), axis=1)
        m = gaussian_blur_u8_1d(m, ks, float(mask_blur), axis=0)
    overlay = np.clip(m.astype(np.int32) * 2, 0, 255).astype(np.uint8)
    return m, overlay


# ---------------------------------------------------------------------------------------------- PIL.Image.resize (BICUBIC)
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_resample_coeffs(in_size: int, out_size: int):
    """Pillow src/libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2)
    over the whole axis: -> (ksize, bounds [out,2] = (xmin, count), int coefficients [out, ksize])."""
    support = 2.0
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)          # left-to-right double accumulation, as the C loop
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bicubic_u8(img: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """Image.fromarray(img).resize((w, h), BICUBIC) for a uint8 single-band image: horizontal pass to uint8, then
    vertical pass (ImagingResampleInner; a pass whose size does not change is skipped)."""
    assert img.dtype == np.uint8 and img.ndim == 2
    H, W = img.shape
    h, w = out_hw
    cur = img
    if w != W:
        _, bounds, kk = pil_resample_coeffs(W, w)
        out = np.zeros((H, w), dtype=np.uint8)
        src = cur.astype(np.int64)
        for xx in range(w):
            x0, n = bounds[xx]
            ss = (src[:, x0:x0 + n] * kk[xx, :n][None, :]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
            out[:, xx] = _clip8(ss)
        cur = out
    if h != H:
        _, bounds, kk = pil_resample_coeffs(H, h)
        out = np.zeros((h, cur.shape[1]), dtype=np.uint8)
        src = cur.astype(np.int64)
        for yy in range(h):
            y0, n = bounds[yy]
            ss = (src[y0:y0 + n, :] * kk[yy, :n][:, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
            out[yy, :] = _clip8(ss)
        cur = out
    return cur


def a1111_latent_mask(blurred_u8: np.ndarray, lat_hw: Tuple[int, int]) -> np.ndarray:
    """latmask = np.around(np.array(mask.convert('RGB').resize((w, h)), float32)[..., 0] / 255) -> {0,1} float32
    (processing.py; 1 = repaint, 0 = keep the original latent)."""
    r = pil_resize_bicubic_u8(blurred_u8, lat_hw)
    return np.around(r.astype(np.float32) / 255.0).astype(np.float32)


# ---------------------------------------------------------------------------------------------- PIL overlay compositing
def _muldiv255(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    t = a * b + 128
    return ((t >> 8) + t) >> 8


def a1111_apply_overlay(generated_u8: np.ndarray, original_u8: np.ndarray, overlay_mask_u8: np.ndarray) -> np.ndarray:
    """processing.py: overlay = Image.new('RGBa').paste(original.convert('RGBA').convert('RGBa'),
    mask=ImageOps.invert(mask_for_overlay)).convert('RGBA'); apply_overlay: generated.convert('RGBA')
    .alpha_composite(overlay).convert('RGB').  Pillow integer arithmetic (Convert.c rgbA2rgba / rgba2rgbA, Paste.c
    paste_mask_L, AlphaComposite.c)."""
    gen = generated_u8.astype(np.int64)
    org = original_u8.astype(np.int64)
    a = 255 - overlay_mask_u8.astype(np.int64)                 # ImageOps.invert
    a3 = a[..., None]
    pre = _muldiv255(org, np.broadcast_to(a3, org.shape))      # premultiplied colour pasted over transparent black
    # rgba2rgbA: un-premultiply (copy at alpha 0 / 255, else CLIP8(255 * c / alpha), integer division)
    safe = np.where(a3 == 0, 1, a3)
    unp = np.where((a3 == 0) | (a3 == 255), pre, np.minimum(255, (255 * pre) // safe))
    # alpha_composite onto an opaque destination: coef1 = a * 128, coef2 = (255 - a) * 128
    tmp = (unp * a3 + gen * (255 - a3)) * 128 + (0x80 << 7)
    out = (((tmp >> 8) + tmp) >> 8) >> 7
    out = np.where(a3 == 0, gen, out)
    return out.astype(np.uint8)


def vae_output_to_u8(x: np.ndarray) -> np.ndarray:
    """A1111 decode post: clamp((x + 1) / 2, 0, 1) * 255 -> astype(uint8) (truncation); x float32 [3,H,W] -> [H,W,3]."""
    v = np.clip((x.astype(np.float32) + np.float32(1.0)) / np.float32(2.0), 0.0, 1.0)
    return (np.float32(255.0) * np.moveaxis(v, 0, 2)).astype(np.uint8)
