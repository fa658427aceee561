"""Oracle: final decode -> PNG bytes (TEST INFRASTRUCTURE ONLY).

generate_data.py:1227  image = image_processor.postprocess(image, output_type="pt", do_denormalize=[True]*B)
                       -> diffusers VaeImageProcessor.denormalize: (images / 2 + 0.5).clamp(0, 1)
generate_data.py:1234  save_image([image[i]], path) -> torchvision.utils.save_image:
                       grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8).numpy(), PIL PNG
Both run in the image's own dtype (fp16 in the reference), i.e. every op rounds to that dtype.
"""
from __future__ import annotations

import io

import numpy as np


def denormalize(images):
    return (images / 2 + 0.5).clamp(0, 1)


def quantize_hwc(image_chw):
    """torchvision.utils.save_image's array for one [C,H,W] image (make_grid of a single image is the image)."""
    import torch
    grid = image_chw.clone()
    return grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8).numpy()


def decode_to_uint8(images, do_denormalize=True) -> np.ndarray:
    """[B,C,H,W] (any float dtype, CPU) -> [B,H,W,C] uint8 exactly as the reference's two calls produce it."""
    x = denormalize(images) if do_denormalize else images
    return np.stack([quantize_hwc(x[i]) for i in range(x.shape[0])])


def png_bytes(hwc_uint8: np.ndarray) -> bytes:
    """PIL's default PNG encoding, what save_image writes (Image.fromarray(ndarr).save(fp, format=None))."""
    from PIL import Image
    buf = io.BytesIO()
    arr = hwc_uint8[:, :, 0] if hwc_uint8.shape[2] == 1 else hwc_uint8
    Image.fromarray(arr).save(buf, format="PNG")
    return buf.getvalue()
