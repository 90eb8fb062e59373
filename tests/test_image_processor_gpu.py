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


@pytest.mark.parametrize("h,w,oh,ow", [(200, 300, 384, 384), (1000, 777, 384, 384), (500, 123, 336, 83), (37, 53, 112, 112),
                                       (640, 480, 448, 336), (384, 384, 384, 384), (1333, 2000, 384, 384)])
def test_device_resize_is_pil_resize(cuda, h, w, oh, ow):
    """mb200_resize_u8_pass x 2 == PIL.Image.resize(BICUBIC) (what the reference's image processors call), bit for bit"""
    from PIL import Image
    from mantis_b200.models.mllava.image_processing import B200ImageProcessor
    rng = np.random.default_rng(52)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    img[: h // 3, : w // 2] = (img[: h // 3, : w // 2] // 128) * 255          # hard edges: ringing must clip like PIL's
    dev = B200ImageProcessor(None, device=cuda)
    got = dev.resize_u8(torch.from_numpy(img).to(cuda), oh, ow).cpu().numpy()
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
    assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.parametrize("kind", ["siglip", "clip"])
def test_device_geometry_end_to_end(cuda, kind):
    """resize (+ centre crop) + rescale + normalise all on the device == PIL resize + the reference's numpy arithmetic"""
    from PIL import Image
    from transformers import CLIPImageProcessor, SiglipImageProcessor
    from mantis_b200.models.mllava.image_processing import B200ImageProcessor
    from oracle.image_oracle import rescale_normalize
    rng = np.random.default_rng(53)
    imgs = [rng.integers(0, 256, size=s, dtype=np.uint8) for s in [(200, 300, 3), (700, 500, 3), (384, 384, 3)]]
    if kind == "siglip":
        ip = SiglipImageProcessor(size={"height": 384, "width": 384})
        ref_u8 = [np.asarray(Image.fromarray(a).resize((384, 384), resample=Image.BICUBIC)) for a in imgs]
    else:
        ip = CLIPImageProcessor(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
        ref_u8 = []
        for a in imgs:
            h, w = a.shape[:2]
            short, long = (w, h) if w <= h else (h, w)
            ns, nl = 336, int(336 * long / short)
            ow, oh = (ns, nl) if w <= h else (nl, ns)
            r = np.asarray(Image.fromarray(a).resize((ow, oh), resample=Image.BICUBIC))
            top, left = (oh - 336) // 2, (ow - 336) // 2
            ref_u8.append(r[top:top + 336, left:left + 336])
    exp = rescale_normalize(np.stack(ref_u8).transpose(0, 3, 1, 2), ip.rescale_factor, ip.image_mean, ip.image_std)
    dev = B200ImageProcessor(ip, device=cuda, dtype=torch.float32, resize_on_device=True)
    got = dev(images=[Image.fromarray(a) for a in imgs])["pixel_values"]
    assert got.shape == exp.shape and np.array_equal(got.cpu().numpy(), exp)
