"""device tail of the image processor == the reference's host pipeline, bit for bit in fp32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["siglip", "clip"])
def test_device_image_processor_matches_host_pipeline(cuda, kind):
    from transformers import CLIPImageProcessor, SiglipImageProcessor
    from mantis_b200.models.mllava.image_processing import B200ImageProcessor
    from oracle.image_oracle import rescale_normalize
    rng = np.random.default_rng(50)
    if kind == "siglip":
        ip = SiglipImageProcessor(size={"height": 384, "width": 384})
    else:
        ip = CLIPImageProcessor(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    imgs = [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for (h, w) in [(200, 300), (384, 384), (500, 123)]]
    imgs[1][:2] = 0; imgs[1][2:4] = 255                                       # extreme pixel values
    dev = B200ImageProcessor(ip, device=cuda, dtype=torch.float32)
    got = dev(images=imgs)["pixel_values"]
    assert got.is_cuda and got.dtype == torch.float32
    raw = np.asarray(ip(images=imgs, return_tensors="np", do_rescale=False, do_normalize=False)["pixel_values"]).astype(np.uint8)
    exp = rescale_normalize(raw, ip.rescale_factor, ip.image_mean, ip.image_std)
    assert got.shape == exp.shape
    assert np.array_equal(got.cpu().numpy(), exp)                            # bit-exact vs the reference's numpy arithmetic
    hf = ip(images=imgs, return_tensors="pt")["pixel_values"]                 # installed transformers (fused torchvision path)
    assert (got.cpu() - hf).abs().max().item() < 1e-6
    bf = B200ImageProcessor(ip, device=cuda, dtype=torch.bfloat16)(images=imgs)["pixel_values"]
    assert bf.dtype == torch.bfloat16 and torch.equal(bf.cpu(), torch.from_numpy(exp).bfloat16())


def test_image_normalize_channels_last_and_ragged_width(ops, cuda):
    from mantis_b200.models.mllava.image_processing import normalization_lut
    rng = np.random.default_rng(51)
    lut = normalization_lut(1 / 255, [0.48, 0.45, 0.40], [0.26, 0.26, 0.27])
    x = rng.integers(0, 256, size=(2, 37, 53, 3), dtype=np.uint8)            # HWC, width not a multiple of 8
    got = ops.image_normalize_u8(torch.from_numpy(x).to(cuda), torch.from_numpy(lut).to(cuda), True, torch.float32)
    exp = np.stack([lut[c][x[..., c]] for c in range(3)], axis=1)
    assert np.array_equal(got.cpu().numpy(), exp)
