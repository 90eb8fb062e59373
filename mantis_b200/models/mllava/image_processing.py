"""Device tail of the image processor (the step before the hot path).

`MLlavaProcessor.__call__` (ref: mantis/models/mllava/processing_llava.py:226-252) hands the PIL images to a HuggingFace
image processor, which resizes (PIL), rescales by 1/255, normalises with mean/std and transposes to CHW -- all on the host in
float32, then the 4-byte/channel result crosses PCIe.  For 32 images of 384x384 per step per GPU that is 57 MB of float
traffic and tens of milliseconds of numpy work per step, comparable to the GPU step itself once that takes ~1.6 s / 4 samples.

B200ImageProcessor keeps the geometry on the host exactly as the wrapped processor does it (so the pixels are the
reference's pixels), ships the resized **uint8** image (4x fewer bytes, pinned) and finishes on the device with one gather
kernel (`mb200_image_normalize_u8`): out[n,c,h,w] = lut[c][pixel].  The 256-entry tables are built on the host with the
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


class B200ImageProcessor:
    """Drop-in for the `image_processor` attribute of MLlavaProcessor: same call signature, `pixel_values` comes back as a
    device tensor [N, 3, H, W] in `dtype` (list inputs of PIL images / numpy arrays, like the wrapped processor)."""

    def __init__(self, image_processor, device="cuda", dtype=torch.bfloat16):
        self.image_processor = image_processor
        self.device = torch.device(device)
        self.dtype = dtype
        self._lut = None
        self._lut_key = None

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

    def preprocess(self, images, return_tensors="pt", do_rescale=None, do_normalize=None, **kwargs):
        ip = self.image_processor
        do_rescale = ip.do_rescale if do_rescale is None else do_rescale
        do_normalize = ip.do_normalize if do_normalize is None else do_normalize
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
