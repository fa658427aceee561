"""Oracle: guided sampling step (TEST INFRASTRUCTURE ONLY).

CPU restatement of generate_data.py:109-137 and :687-767 with the module-global
``args`` made an explicit parameter and the hard-coded ``.cuda()`` calls dropped
(the reference has no CPU path: generate_data.py:692,695).  Everything else --
operation order, where features are / are not normalised, the in-place masked
clamp, ``score / guidance_period``, the SGD step on the channel-affine params --
follows the reference line by line so that ``torch.autograd`` on CPU produces
the reference's gradients.

``unet``, ``vae``, ``image_encoder`` and ``image_processor`` are duck-typed
callables so tiny random-init stand-ins can be plugged in:
    unet(x, t, prompt_embeds, class_labels=None, return_dict=False)[0]
    vae.decode(z, return_dict=False, generator=None)[0] ; vae.config.scaling_factor
    image_encoder.encode_image(img)
    image_processor.postprocess(img, output_type="pt", do_denormalize=[...])
"""
from __future__ import annotations

import torch

from . import energy as _energy


def denoise_one_step(args, latents, noise_scheduler, t, unet, prompt_embeds, class_labels):
    """generate_data.py:109-121."""
    latent_model_input = torch.cat([latents] * 2) if args.do_classifier_free_guidance else latents
    latent_model_input = noise_scheduler.scale_model_input(latent_model_input, t)
    noise_pred = unet(latent_model_input, t, prompt_embeds, class_labels=class_labels, return_dict=False)[0]
    if args.do_classifier_free_guidance:
        noise_pred_uncond, noise_pred_text = noise_pred.chunk(2)
        noise_pred = noise_pred_uncond + args.guidance_scale * (noise_pred_text - noise_pred_uncond)
    ddim_output = noise_scheduler.step(noise_pred, t, latents, return_dict=True)
    return ddim_output["prev_sample"], ddim_output["pred_original_sample"]


def tensor_clamp(t, min, max, in_place=True):
    """generate_data.py:124-133 (boolean-mask assignment on .data)."""
    res = t if in_place else t.clone()
    idx = res.data < min
    res.data[idx] = min[idx]
    idx = res.data > max
    res.data[idx] = max[idx]
    return res


def linfball_proj(center, radius, t, in_place=True):
    """generate_data.py:136-137."""
    return tensor_clamp(t, min=center - radius, max=center + radius, in_place=in_place)


def _features(args, pred_x0, vae, image_encoder, image_processor, generator):
    """generate_data.py:701-705 / 743-746: decode -> postprocess(identity) -> bicubic 224 -> guide."""
    D_x0_t = vae.decode(pred_x0 / vae.config.scaling_factor, return_dict=False, generator=generator)[0]
    D_x0_t = image_processor.postprocess(D_x0_t, output_type="pt", do_denormalize=[False] * D_x0_t.shape[0])
    D_x0_t = torch.nn.functional.interpolate(D_x0_t, size=(224, 224), mode="bicubic")
    return image_encoder.encode_image(D_x0_t).float()


def transform_guidance(args, latents, batch, sub_timesteps, noise_scheduler, unet, prompt_embeds, class_labels,
                       vae, image_encoder, image_processor, weight_dtype, generator,
                       total_global_proto, total_local_proto, channel_noise=None, channel_noise_bias=None):
    """generate_data.py:687-732.  ``channel_noise``/``channel_noise_bias`` default to the reference's
    CPU-global-RNG draws (:692-695) and may be passed in to pin them."""
    bs = latents.shape[0]
    channel_noise_dim = latents.shape[1]
    if channel_noise is None:
        channel_noise = torch.rand([bs, channel_noise_dim, 1, 1])
    if channel_noise_bias is None:
        channel_noise_bias = torch.zeros([bs, channel_noise_dim, 1, 1]).normal_(0, 1)
    channel_noise = channel_noise.clone().to(dtype=weight_dtype).requires_grad_(True)
    channel_noise_bias = channel_noise_bias.clone().to(dtype=weight_dtype).requires_grad_(True)
    x_dec_noisy = latents * (1 + channel_noise) + channel_noise_bias

    score = 0.0
    for temp_t in sub_timesteps:
        x_dec_noisy, pred_x0 = denoise_one_step(args, x_dec_noisy, noise_scheduler, temp_t, unet, prompt_embeds,
                                                class_labels)
        image_features = _features(args, pred_x0, vae, image_encoder, image_processor, generator)
        score = score + _energy.energy_score(image_features, batch["targets"], total_global_proto,
                                             total_local_proto, args.gs, args.ls, normalize_f=False)
    score = score / args.guidance_period

    channel_noise_grad, channel_noise_bias_grad = torch.autograd.grad(score, [channel_noise, channel_noise_bias])
    channel_noise.data.add_(-args.rho * channel_noise_grad)
    channel_noise_bias.data.add_(-args.rho * channel_noise_bias_grad)

    x_dec_temp = latents.clone()
    latents = latents * (1 + channel_noise) + channel_noise_bias
    linfball_proj(x_dec_temp, args.constraint_value, latents, in_place=True)
    return latents.detach(), score


def direct_guidance(args, latents, batch, t_i, noise_scheduler, unet, prompt_embeds, class_labels,
                    vae, image_encoder, image_processor, weight_dtype, generator,
                    total_global_proto, total_local_proto):
    """generate_data.py:735-767."""
    latents = latents.detach().clone().requires_grad_(True)
    x_dec_next, x_0 = denoise_one_step(args, latents, noise_scheduler, t_i, unet, prompt_embeds, class_labels)
    image_features = _features(args, x_0, vae, image_encoder, image_processor, generator)
    score = _energy.energy_score(image_features, batch["targets"], total_global_proto, total_local_proto,
                                 args.gs, args.ls, normalize_f=True)
    x_dec_grad = torch.autograd.grad(score, latents)[0]
    x_dec_next = x_dec_next - args.rho * x_dec_grad
    return x_dec_next.detach(), x_0.detach(), score


def affine_project(x, a, b, radius=None, center=None):
    """Closed form of generate_data.py:696 / :726-728: y = x*(1+a)+b, optionally clamped to center +- radius."""
    y = x * (1 + a) + b
    if radius is not None and radius >= 0:
        c = x if center is None else center
        y = torch.maximum(torch.minimum(y, c + radius), c - radius)
    return y
