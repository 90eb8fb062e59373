"""TEST INFRASTRUCTURE ONLY (see oracle/ref_shim.py): numpy restatement of the image processor tail the reference runs on
the host -- `rescale` then `normalize` of transformers/image_transforms.py (the numpy "slow" image processors that the
reference's `transformers<4.46` pin uses for SiglipImageProcessor / CLIPImageProcessor, called from
mantis/models/mllava/processing_llava.py:226-252), followed by the HWC -> CHW transpose.

    rescale  : image.astype(np.float64) * scale, cast to float32
    normalize: (image - mean) / std with mean / std cast to the image dtype (float32)

Pinned against the installed transformers (5.5, torchvision backend, which fuses the two steps) to 1e-6 in
tests/test_image_processor_gpu.py -- parity unpinned by reference-owned vectors (the reference has no tests)."""
import numpy as np


def rescale_normalize(pixels_u8_nchw: np.ndarray, rescale_factor: float, mean, std) -> np.ndarray:
    x = (pixels_u8_nchw.astype(np.float64) * rescale_factor).astype(np.float32)
    m = np.asarray(mean, dtype=np.float32).reshape(1, -1, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(1, -1, 1, 1)
    return (x - m) / s
