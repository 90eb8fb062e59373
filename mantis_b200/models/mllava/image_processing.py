"""Device tail of the image processor (the step before the hot path).

`MLlavaProcessor.__call__` (ref: mantis/models/mllava/processing_llava.py:226-252) hands the PIL images to a HuggingFace
image processor, which resizes (PIL), rescales by 1/255, normalises with mean/std and transposes to CHW -- all on the host in
float32, then the 4-byte/channel result crosses PCIe.  For 32 images of 384x384 per step per GPU that is 57 MB of float
traffic and tens of milliseconds of numpy work per step, comparable to the GPU step itself once that takes ~1.6 s / 4 samples.

B200ImageProcessor by default keeps the geometry on the host exactly as the wrapped processor does it (so the pixels are
the reference's pixels), ships the resized **uint8** image (4x fewer bytes, pinned) and finishes on the device with one
gather kernel (`mb200_image_normalize_u8`): out[n,c,h,w] = lut[c][pixel].  With `resize_on_device=True` the bicubic resize
moves to the device as well (`mb200_resize_u8_pass`, Pillow's two-pass 22-bit fixed-point resampler with host-computed tap
tables: bit-identical to `PIL.Image.resize`, which is what the reference's processors call) for the two geometries on the
path: SigLIP's fixed `height x width` resize and CLIP's shortest-edge resize + centre crop.  The 256-entry tables are built on the host with the
reference's own numpy arithmetic -- float32(float64(v) * rescale_factor), then (x - mean) / std in float32
(transformers/image_transforms.py `rescale` / `normalize`, the slow path the reference's transformers<4.46 pin uses) -- so
the device result is bit-identical to the reference's for every possible pixel value.
"""
import numpy as np
import torch
from transformers.feature_extraction_utils import BatchFeature

from ... import ops


def normalization_lut(rescale_factor, image_mean, image_std, do_rescale=True, do_normalize=True) -> np.ndarray:
    """[C, 256] float32 table: what the reference's numpy pipeline yields for pixel value v in channel c."""
    v = np.arange(256, dtype=np.uint8)
    x = v.astype(np.float64) * rescale_factor if do_rescale else v.astype(np.float64)
    x = x.astype(np.float32)
    C = len(image_mean) if do_normalize else 3
    out = np.empty((C, 256), dtype=np.float32)
    for c in range(C):
        if do_normalize:
            out[c] = (x - np.float32(image_mean[c])) / np.float32(image_std[c])
        else:
            out[c] = x
    return out


_PRECISION_BITS = 32 - 8 - 2
_coeff_cache = {}


def bicubic_coeffs(in_size: int, out_size: int):
    """Tap ranges and 22-bit fixed-point weights of Pillow's antialiased bicubic resampler (src/libImaging/Resample.c:
    precompute_coeffs + normalize_coeffs_8bpc) for one axis: (bounds int32 [out, 2] = (first tap, taps), coef int32
    [out, ksize]).  Every floating-point step is done in double precision in Pillow's operation order (the weight sum is
    accumulated tap by tap), so the integers are Pillow's integers."""
    key = (int(in_size), int(out_size))
    hit = _coeff_cache.get(key)
    if hit is not None:
        return hit
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale                                   # bicubic support = 2
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)
    xmax = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size)
    count = xmax - xmin
    taps = np.arange(ksize, dtype=np.int64)[None, :]
    x = ((taps + xmin[:, None]) - center[:, None] + 0.5) * ss
    x = np.abs(x)
    a = -0.5
    w = np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1, np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))
    w = np.where(taps < count[:, None], w, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for t in range(ksize):                                        # sequential sum, like the C loop
        ww = ww + w[:, t]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    scaled = w * float(1 << _PRECISION_BITS)
    coef = np.where(w < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int32)
    bounds = np.stack([xmin, count], axis=1).astype(np.int32)
    if len(_coeff_cache) > 256:
        _coeff_cache.clear()
    _coeff_cache[key] = (bounds, coef)
    return bounds, coef


class B200ImageProcessor:
    """Drop-in for the `image_processor` attribute of MLlavaProcessor: same call signature, `pixel_values` comes back as a
    device tensor [N, 3, H, W] in `dtype` (list inputs of PIL images / numpy arrays, like the wrapped processor)."""

    def __init__(self, image_processor, device="cuda", dtype=torch.bfloat16, resize_on_device=False):
        self.image_processor = image_processor
        self.device = torch.device(device)
        self.dtype = dtype
        self.resize_on_device = resize_on_device
        self._lut = None
        self._lut_key = None
        self._dev_coeffs = {}

    def __getattr__(self, name):                      # size / crop_size / image_mean ... are read by callers
        return getattr(self.__dict__["image_processor"], name)

    def _table(self, do_rescale, do_normalize):
        ip = self.image_processor
        key = (bool(do_rescale), bool(do_normalize), float(ip.rescale_factor), tuple(ip.image_mean), tuple(ip.image_std))
        if key != self._lut_key:
            lut = normalization_lut(ip.rescale_factor, ip.image_mean, ip.image_std, do_rescale, do_normalize)
            self._lut = torch.from_numpy(lut).to(self.device)
            self._lut_key = key
        return self._lut

    # ---- geometry on the device (optional) ----
    def _coeffs(self, in_size, out_size):
        key = (in_size, out_size)
        hit = self._dev_coeffs.get(key)
        if hit is None:
            b, c = bicubic_coeffs(in_size, out_size)
            hit = (torch.from_numpy(b).to(self.device), torch.from_numpy(c).to(self.device))
            if len(self._dev_coeffs) > 256:
                self._dev_coeffs.clear()
            self._dev_coeffs[key] = hit
        return hit

    def resize_u8(self, img, out_h, out_w):
        """[H, W, C] uint8 device tensor -> [out_h, out_w, C], == PIL.Image.resize((out_w, out_h), BICUBIC)"""
        H, W, _ = img.shape
        if W != out_w:
            img = ops.resize_u8_pass(img, *self._coeffs(W, out_w), out_w, horizontal=True)
        if H != out_h:
            img = ops.resize_u8_pass(img, *self._coeffs(H, out_h), out_h, horizontal=False)
        return img

    def _device_geometry(self, images):
        """-> uint8 [N, H, W, C] device tensor, or None when the wrapped processor's geometry is not one of the two on the
        path (then the host does it).  SigLIP: resize to size[height] x size[width]; CLIP: shortest edge -> size, longer edge
        int(size * long / short), then the centre crop_size window."""
        ip = self.image_processor
        size = dict(getattr(ip, "size", None) or {})
        if not getattr(ip, "do_resize", True) or int(getattr(ip, "resample", 3)) != 3:       # 3 = PIL BICUBIC
            return None
        crop = dict(getattr(ip, "crop_size", None) or {}) if getattr(ip, "do_center_crop", False) else None
        outs = []
        for im in images:
            arr = np.asarray(im.convert("RGB")) if hasattr(im, "convert") else np.asarray(im)
            if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
                return None
            h, w = arr.shape[:2]
            if "height" in size and "width" in size:
                oh, ow = int(size["height"]), int(size["width"])
            elif "shortest_edge" in size:
                s0 = int(size["shortest_edge"])
                short, long = (w, h) if w <= h else (h, w)
                new_short, new_long = s0, int(s0 * long / short)
                ow, oh = (new_short, new_long) if w <= h else (new_long, new_short)
            else:
                return None
            t = torch.from_numpy(np.array(arr, copy=True))          # PIL hands out read-only buffers
            if self.device.type == "cuda":
                t = t.pin_memory()
            t = self.resize_u8(t.to(self.device, non_blocking=True), oh, ow)
            if crop is not None:
                ch, cw = int(crop["height"]), int(crop["width"])
                if ch > oh or cw > ow:
                    return None                                   # HF pads here; leave that corner to the host path
                top, left = (oh - ch) // 2, (ow - cw) // 2
                t = t[top:top + ch, left:left + cw]
            outs.append(t)
        if len({tuple(t.shape) for t in outs}) != 1:
            return None
        return torch.stack(outs, dim=0)

    def preprocess(self, images, return_tensors="pt", do_rescale=None, do_normalize=None, **kwargs):
        ip = self.image_processor
        do_rescale = ip.do_rescale if do_rescale is None else do_rescale
        do_normalize = ip.do_normalize if do_normalize is None else do_normalize
        if self.resize_on_device and not kwargs:
            imgs = images if isinstance(images, (list, tuple)) else [images]
            px = self._device_geometry(imgs)
            if px is not None:
                out = ops.image_normalize_u8(px, self._table(do_rescale, do_normalize), channels_last=True, out_dtype=self.dtype)
                return BatchFeature(data={"pixel_values": out})
        raw = ip(images=images, return_tensors="np", do_rescale=False, do_normalize=False, **kwargs)["pixel_values"]
        raw = np.ascontiguousarray(np.asarray(raw))
        if raw.dtype != np.uint8:
            if np.abs(raw - np.rint(raw)).max() != 0 or raw.min() < 0 or raw.max() > 255:
                raise ValueError("the wrapped image processor did not return 8-bit pixel values")
            raw = raw.astype(np.uint8)
        host = torch.from_numpy(raw)
        if self.device.type == "cuda":
            host = host.pin_memory()
        px = host.to(self.device, non_blocking=True)                       # uint8 [N, C, H, W]
        out = ops.image_normalize_u8(px, self._table(do_rescale, do_normalize), channels_last=False, out_dtype=self.dtype)
        return BatchFeature(data={"pixel_values": out})

    __call__ = preprocess
