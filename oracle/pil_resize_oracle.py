"""TEST INFRASTRUCTURE ONLY: numpy restatement of Pillow's antialiased resize for 8-bit images (`Image.resize(size,
resample=BICUBIC)`), which is what the reference's image processors call for every image
(mantis/models/mllava/processing_llava.py:226-252 -> transformers image_transforms.resize -> PIL).

Third-party algorithm (Pillow, src/libImaging/Resample.c; the image ships Pillow and the tests pin this restatement to
PIL.Image.resize itself, bit for bit):
  precompute_coeffs   per output index: centre = (xx + 0.5) * scale, support = 2 * max(scale, 1) for the bicubic filter
                      (a = -0.5), taps xmin .. xmax-1 clipped to the image, weights normalised to sum 1 in double precision
  normalize_coeffs_8bpc   weights -> 22-bit fixed point, rounded half away from zero
  ImagingResampleHorizontal_8bpc / Vertical_8bpc   acc = 2^21 + sum(pixel * k); out = clip8(acc >> 22); horizontal pass
                      first (uint8 intermediate), then vertical
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int, support_base: float = 2.0, filt=_bicubic):
    """-> (bounds int32 [out, 2] = (xmin, count), coeffs int32 [out, ksize]) exactly as Pillow computes them"""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support_base * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            w = filt((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            for x in range(xmax):
                kk[xx, x] /= ww
        bounds[xx] = (xmin, xmax)
    scaled = kk * float(1 << PRECISION_BITS)
    ki = np.where(kk < 0, (-0.5 + scaled).astype(np.int64), (0.5 + scaled).astype(np.int64)).astype(np.int32)   # C (int) truncates
    return bounds, ki


def _pass(img: np.ndarray, bounds, ki, axis: int) -> np.ndarray:
    """one resampling pass over `axis` (0 = vertical, 1 = horizontal) of an [H, W, C] uint8 image"""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(ki[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """[H, W, C] uint8 -> [out_h, out_w, C] uint8, == np.asarray(PIL.Image.fromarray(img).resize((out_w, out_h), BICUBIC))"""
    h, w = img.shape[:2]
    if w != out_w:
        img = _pass(img, *precompute_coeffs(w, out_w), axis=1)
    if h != out_h:
        img = _pass(img, *precompute_coeffs(h, out_h), axis=0)
    return img
