"""SD-v1.x-shaped UNet / VAE and a ResNet-50 guide in plain PyTorch (diffusers / timm are not installable here).

These are NOT part of the hot path being accelerated -- north_star keeps "the UNet, VAE and guide-network
forward/backward on PyTorch's own kernels".  They exist so that ``generate_data.py`` and ``bench.py`` can run
the reference's configuration (SD v1.4 shapes, random-init weights, no network) end to end.  Call
signatures are the diffusers / reference ones:

    unet(sample, timestep, encoder_hidden_states, class_labels=None, return_dict=False)[0]   generate_data.py:112
    vae.decode(z, return_dict=False, generator=None)[0] ; vae.config.scaling_factor          :701, :1223
    vae.encode(x).latent_dist.sample()                                                        dataloader.py:808
    image_encoder.encode_image(x, pooling='avg')                                              model_utils.py:29-41
"""
from __future__ import annotations

import math
import types

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint


# --------------------------------------------------------------------------------------------- UNet
class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch=None, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout) if temb_ch else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None and temb is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Attention(nn.Module):
    def __init__(self, dim, ctx_dim=None, heads=8):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_out = nn.Linear(dim, dim)

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, N, Cd = x.shape
        q = self.to_q(x).view(B, N, self.heads, -1).transpose(1, 2)
        k = self.to_k(ctx).view(B, ctx.shape[1], self.heads, -1).transpose(1, 2)
        v = self.to_v(ctx).view(B, ctx.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.to_out(o.transpose(1, 2).reshape(B, N, Cd))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, ctx_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff_in = nn.Linear(dim, dim * 8)  # GEGLU
        self.ff_out = nn.Linear(dim * 4, dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        h, gate = self.ff_in(self.norm3(x)).chunk(2, dim=-1)
        return x + self.ff_out(h * F.gelu(gate))


class Transformer2D(nn.Module):
    def __init__(self, ch, ctx_dim, heads):
        super().__init__()
        self.norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.proj_in = nn.Conv2d(ch, ch, 1)
        self.block = BasicTransformerBlock(ch, ctx_dim, heads)
        self.proj_out = nn.Conv2d(ch, ch, 1)

    def forward(self, x, ctx):
        B, Cc, H, W = x.shape
        h = self.proj_in(self.norm(x)).flatten(2).transpose(1, 2)
        h = self.block(h, ctx)
        return x + self.proj_out(h.transpose(1, 2).reshape(B, Cc, H, W))


def timestep_embedding(t, dim, dtype):
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1).to(dtype)  # flip_sin_to_cos=True


class UNet2DConditionModel(nn.Module):
    """SD v1.x: in/out 4, channels (320,640,1280,1280), 2 res layers per block, 8 heads, cross-attn dim 768."""

    def __init__(self, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 cross_attention_dim=768, heads=8, sample_size=64):
        super().__init__()
        self.config = types.SimpleNamespace(in_channels=in_channels, out_channels=out_channels, sample_size=sample_size,
                                            cross_attention_dim=cross_attention_dim)
        self.sample_size = sample_size
        self.gradient_checkpointing = False
        ch0 = block_out_channels[0]
        temb = ch0 * 4
        self.time_embedding = nn.Sequential(nn.Linear(ch0, temb), nn.SiLU(), nn.Linear(temb, temb))
        self.conv_in = nn.Conv2d(in_channels, ch0, 3, padding=1)
        self.down = nn.ModuleList()
        skip_chs = [ch0]
        cin = ch0
        n = len(block_out_channels)
        for i, cout in enumerate(block_out_channels):
            has_attn = i < n - 1
            blk = nn.ModuleDict(dict(res=nn.ModuleList(), attn=nn.ModuleList()))
            for _ in range(layers_per_block):
                blk["res"].append(ResnetBlock2D(cin, cout, temb))
                if has_attn:
                    blk["attn"].append(Transformer2D(cout, cross_attention_dim, heads))
                cin = cout
                skip_chs.append(cout)
            if i < n - 1:
                blk["down"] = nn.Conv2d(cout, cout, 3, stride=2, padding=1)
                skip_chs.append(cout)
            self.down.append(blk)
        self.mid_res1 = ResnetBlock2D(cin, cin, temb)
        self.mid_attn = Transformer2D(cin, cross_attention_dim, heads)
        self.mid_res2 = ResnetBlock2D(cin, cin, temb)
        self.up = nn.ModuleList()
        rev = list(reversed(block_out_channels))
        for i, cout in enumerate(rev):
            has_attn = i > 0
            blk = nn.ModuleDict(dict(res=nn.ModuleList(), attn=nn.ModuleList()))
            for _ in range(layers_per_block + 1):
                blk["res"].append(ResnetBlock2D(cin + skip_chs.pop(), cout, temb))
                if has_attn:
                    blk["attn"].append(Transformer2D(cout, cross_attention_dim, heads))
                cin = cout
            if i < n - 1:
                blk["up"] = nn.Conv2d(cout, cout, 3, padding=1)
            self.up.append(blk)
        self.conv_norm_out = nn.GroupNorm(32, ch0)
        self.conv_out = nn.Conv2d(ch0, out_channels, 3, padding=1)
        self._ch0 = ch0

    def enable_gradient_checkpointing(self):  # generate_data.py:1049-1050
        self.gradient_checkpointing = True

    def _run(self, fn, *a):
        if self.gradient_checkpointing and torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in a):
            return checkpoint(fn, *a, use_reentrant=False)
        return fn(*a)

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, return_dict=False):
        t = getattr(timestep, "dev", None)       # expand._DevTimestep: an int that already carries its device tensor (graph capture)
        if t is None:
            t = torch.as_tensor(timestep, device=sample.device)
        t = t.reshape(-1)
        if t.numel() == 1:
            t = t.expand(sample.shape[0])
        temb = self.time_embedding(timestep_embedding(t, self._ch0, sample.dtype))
        ctx = encoder_hidden_states
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down:
            for j, res in enumerate(blk["res"]):
                h = self._run(res, h, temb)
                if len(blk["attn"]):
                    h = self._run(blk["attn"][j], h, ctx)
                skips.append(h)
            if "down" in blk:
                h = blk["down"](h)
                skips.append(h)
        h = self._run(self.mid_res1, h, temb)
        h = self._run(self.mid_attn, h, ctx)
        h = self._run(self.mid_res2, h, temb)
        for blk in self.up:
            for j, res in enumerate(blk["res"]):
                h = self._run(res, torch.cat([h, skips.pop()], dim=1), temb)
                if len(blk["attn"]):
                    h = self._run(blk["attn"][j], h, ctx)
            if "up" in blk:
                h = blk["up"](F.interpolate(h, scale_factor=2.0, mode="nearest"))
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return (h,)


# --------------------------------------------------------------------------------------------- VAE
class VaeAttention(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.attn = Attention(ch, None, heads=1)

    def forward(self, x):
        B, Cc, H, W = x.shape
        h = self.group_norm(x).flatten(2).transpose(1, 2)
        return x + self.attn(h).transpose(1, 2).reshape(B, Cc, H, W)


class VaeDecoder(nn.Module):
    def __init__(self, latent=4, chs=(128, 256, 512, 512), layers=2):
        super().__init__()
        top = chs[-1]
        self.conv_in = nn.Conv2d(latent, top, 3, padding=1)
        self.mid = nn.ModuleList([ResnetBlock2D(top, top, None, eps=1e-6), VaeAttention(top), ResnetBlock2D(top, top, None, eps=1e-6)])
        self.up = nn.ModuleList()
        cin = top
        for i, cout in enumerate(reversed(chs)):
            blk = nn.ModuleDict(dict(res=nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, None, eps=1e-6)
                                                        for j in range(layers + 1)])))
            if i < len(chs) - 1:
                blk["up"] = nn.Conv2d(cout, cout, 3, padding=1)
            self.up.append(blk)
            cin = cout
        self.conv_norm_out = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv_out = nn.Conv2d(cin, 3, 3, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        for m in self.mid:
            h = m(h)
        for blk in self.up:
            for res in blk["res"]:
                h = res(h)
            if "up" in blk:
                h = blk["up"](F.interpolate(h, scale_factor=2.0, mode="nearest"))
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class VaeEncoder(nn.Module):
    def __init__(self, latent=4, chs=(128, 256, 512, 512), layers=2):
        super().__init__()
        self.conv_in = nn.Conv2d(3, chs[0], 3, padding=1)
        self.down = nn.ModuleList()
        cin = chs[0]
        for i, cout in enumerate(chs):
            blk = nn.ModuleDict(dict(res=nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, None, eps=1e-6)
                                                        for j in range(layers)])))
            if i < len(chs) - 1:
                blk["down"] = nn.Conv2d(cout, cout, 3, stride=2, padding=0)
            self.down.append(blk)
            cin = cout
        self.mid = nn.ModuleList([ResnetBlock2D(cin, cin, None, eps=1e-6), VaeAttention(cin), ResnetBlock2D(cin, cin, None, eps=1e-6)])
        self.conv_norm_out = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv_out = nn.Conv2d(cin, 2 * latent, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for blk in self.down:
            for res in blk["res"]:
                h = res(h)
            if "down" in blk:
                h = blk["down"](F.pad(h, (0, 1, 0, 1)))
        for m in self.mid:
            h = m(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class _DiagonalGaussian:
    def __init__(self, moments):
        self.mean, logvar = moments.chunk(2, dim=1)
        self.std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))

    def sample(self, generator=None):
        return self.mean + self.std * torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)

    def mode(self):
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, latent=4, chs=(128, 256, 512, 512), scaling_factor=0.18215, with_encoder=True):
        super().__init__()
        self.config = types.SimpleNamespace(scaling_factor=scaling_factor, latent_channels=latent)
        self.encoder = VaeEncoder(latent, chs) if with_encoder else None
        self.quant_conv = nn.Conv2d(2 * latent, 2 * latent, 1) if with_encoder else None
        self.post_quant_conv = nn.Conv2d(latent, latent, 1)
        self.decoder = VaeDecoder(latent, chs)

    def encode(self, x):
        return types.SimpleNamespace(latent_dist=_DiagonalGaussian(self.quant_conv(self.encoder(x))))

    def decode(self, z, return_dict=False, generator=None):
        return (self.decoder(self.post_quant_conv(z)),)


class VaeImageProcessor:
    """diffusers.VaeImageProcessor.postprocess(output_type='pt') (generate_data.py:703,744,1227)."""

    def postprocess(self, image, output_type="pt", do_denormalize=None):
        if do_denormalize is None or not any(do_denormalize):
            return image
        if all(do_denormalize):
            return (image / 2 + 0.5).clamp(0, 1)
        return torch.stack([(im / 2 + 0.5).clamp(0, 1) if d else im for im, d in zip(image, do_denormalize)])


# --------------------------------------------------------------------------------------------- guide
def add_encoder_image_method(model):
    """model_utils.py:29-41 -- forward_features -> AdaptiveAvgPool2d(1) -> flatten."""

    def encode_image(self, x, pooling="avg"):
        f = self.forward_features(x)
        if pooling == "avg":
            f = F.adaptive_avg_pool2d(f, 1).flatten(1)
        return f

    model.encode_image = types.MethodType(encode_image, model)
    return model


def create_model(model_name="resnet50", num_classes=100, pretrained=False, class_names=None, cache_dir=None,
                 dataset_name=None, weight_path=None):
    """model_utils.py:43-104 for the ResNet family, on torchvision (same parameter names as timm's resnet50, so a
    reference ``model_best.pth.tar`` loads after the ``module.`` strip, model_utils.py:89-101)."""
    import torchvision
    ctor = {"resnet50": torchvision.models.resnet50, "resnext50": torchvision.models.resnext50_32x4d,
            "wideresnet50": torchvision.models.wide_resnet50_2, "resnet18": torchvision.models.resnet18}
    if model_name not in ctor:
        raise ValueError(f"guide arch {model_name!r} is not available offline (have: {sorted(ctor)})")
    model = ctor[model_name](weights=None, num_classes=num_classes)

    def forward_features(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))

    model.forward_features = types.MethodType(forward_features, model)
    add_encoder_image_method(model)
    if weight_path is not None:
        if not os.path.exists(weight_path):   # the reference's torch.load would raise too (model_utils.py:89); never fall back to random init
            raise FileNotFoundError(f"guide weights {weight_path!r} not found (--encoder_weight_path)")
        ck = torch.load(weight_path, map_location="cpu", weights_only=False)
        sd = ck.get("state_dict", ck)
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
        model.load_state_dict(sd)
    return model


import os  # noqa: E402  (used by create_model)
