"""Tiny random-init stand-ins for UNet / VAE / guide encoder / image processor.

Shared by tests/golden/make_golden.py (which runs the reference's own function bodies on them)
and by the parity tests.  Duck-typed to exactly what generate_data.py calls:
    unet(x, t, prompt_embeds, class_labels=None, return_dict=False)[0]     generate_data.py:112
    vae.decode(z, return_dict=False, generator=None)[0]; vae.config.scaling_factor   :701
    image_processor.postprocess(img, output_type="pt", do_denormalize=[...])          :703
    image_encoder.encode_image(img)                                                   :705
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn
import torch.nn.functional as F


class TinyUNet(nn.Module):
    def __init__(self, ch=4, hidden=16, ctx=8):
        super().__init__()
        self.conv1 = nn.Conv2d(ch, hidden, 3, padding=1)
        self.temb = nn.Linear(1, hidden)
        self.ctx = nn.Linear(ctx, hidden)
        self.conv2 = nn.Conv2d(hidden, ch, 3, padding=1)

    def forward(self, x, t, prompt_embeds, class_labels=None, return_dict=False):
        tt = torch.as_tensor(t, dtype=x.dtype, device=x.device).reshape(1, 1) / 1000.0
        h = self.conv1(x) + self.temb(tt)[:, :, None, None] + self.ctx(prompt_embeds.mean(1))[:, :, None, None]
        return (self.conv2(F.silu(h)),)


class TinyVAE(nn.Module):
    def __init__(self, ch=4, hidden=8, up=8):
        super().__init__()
        self.config = types.SimpleNamespace(scaling_factor=0.18215)
        self.up = up
        self.conv1 = nn.Conv2d(ch, hidden, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden, 3, 3, padding=1)

    def decode(self, z, return_dict=False, generator=None):
        h = F.silu(self.conv1(z))
        h = F.interpolate(h, scale_factor=self.up, mode="nearest")
        return (torch.tanh(self.conv2(h)),)


class TinyEncoder(nn.Module):
    def __init__(self, dim=64, hidden=8):
        super().__init__()
        self.conv = nn.Conv2d(3, hidden, 8, stride=8)
        self.fc = nn.Linear(hidden, dim)

    def encode_image(self, x, pooling="avg"):
        h = F.silu(self.conv(x)).mean((2, 3))
        return self.fc(h)


class IdentityProcessor:
    """VaeImageProcessor.postprocess(output_type='pt'): identity, or (x/2+0.5).clamp(0,1) when denormalising."""

    def postprocess(self, image, output_type="pt", do_denormalize=None):
        if do_denormalize is None or not any(do_denormalize):
            return image
        return torch.stack([(im / 2 + 0.5).clamp(0, 1) if d else im for im, d in zip(image, do_denormalize)])


def make_nets(seed=0, feat_dim=64, ctx=8):
    g = torch.Generator().manual_seed(seed)
    nets = dict(unet=TinyUNet(ctx=ctx), vae=TinyVAE(), enc=TinyEncoder(dim=feat_dim))
    for m in nets.values():
        for p in m.parameters():
            with torch.no_grad():
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 / max(1.0, p[0].numel() ** 0.5) if p.dim() > 1 else 0.1))
        m.eval()
    return nets["unet"], nets["vae"], nets["enc"]
