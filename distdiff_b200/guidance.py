"""Drop-in mirror of the reference's guided sampling step (generate_data.py:109-137, 687-767).

Same function names, argument order and return values as the reference so that its ``main()`` loop
(:1199-1218) runs unchanged on top of this module.  Like the reference the functions read a module-global
``args`` (set it with ``set_args``).  What changed underneath: every tensor op of the step that is not the
UNet / VAE / guide network is one fused sm_100a kernel (``distdiff_b200.ops``):

    reference eager sequence                               here
    cat/chunk, sub, mul, add, DDIMScheduler.step (~13)   -> K5  dd_cfg_ddim_fwd (+ dd_cfg_ddim_bwd under autograd)
    latents*(1+a)+b (3) and its autograd reduce          -> K6  dd_affine_project_fwd / dd_affine_bwd
    gathers, 2 norms, bmm, argmax, index, means (~14+14) -> K4  dd_energy_fwd_bwd (forward AND analytic gradient)
    F.interpolate(bicubic, 224) + its atomic backward    -> K8  dd_bicubic_resize_fwd / _bwd (deterministic gather)
    transform + tensor_clamp masks/scatters (~9)         -> K6  dd_affine_project_fwd (radius >= 0)
    x_next - rho*grad (2)                                -> fused into K5's epilogue

RNG parity: the channel noise is drawn exactly like the reference -- torch.rand then normal_ on the CPU global
generator, in that order (:692-695) -- and moved to the device.
"""
from __future__ import annotations

import torch

from . import ops

args = None  # the reference's module-global argparse namespace (generate_data.py:1253)


def set_args(namespace) -> None:
    global args
    args = namespace


def _need_args():
    if args is None:
        raise RuntimeError("distdiff_b200.guidance.args is not set (call set_args(args) like generate_data.main does)")
    return args


def denoise_one_step(latents, noise_scheduler, t, unet, prompt_embeds, class_labels):
    """generate_data.py:109-121 -> (prev_sample, pred_original_sample)."""
    a = _need_args()
    cfg = bool(a.do_classifier_free_guidance)
    latent_model_input = torch.cat([latents] * 2) if cfg else latents
    latent_model_input = noise_scheduler.scale_model_input(latent_model_input, t)
    noise_pred = unet(latent_model_input, t, prompt_embeds, class_labels=class_labels, return_dict=False)[0]
    a_t, a_prev = noise_scheduler.alpha_pair(t)
    if torch.is_grad_enabled() and (noise_pred.requires_grad or latents.requires_grad):
        return ops.CfgDdimStep.apply(noise_pred, latents, float(a.guidance_scale), a_t, a_prev, cfg)
    return ops.cfg_ddim_step(noise_pred, latents, float(a.guidance_scale), a_t, a_prev, cfg=cfg)


def linfball_proj(center, radius, t, in_place=True):
    """generate_data.py:136-137 (tensor_clamp(t, center - radius, center + radius)) as one K6 launch."""
    B, Cc = t.shape[0], t.shape[1]
    zero = torch.zeros(B * Cc, dtype=torch.float32, device=t.device)
    res = ops.affine_project(t.data, zero, zero, float(radius), center=center)
    if in_place:
        t.data.copy_(res)
        return t
    return res


def _guide_features(pred_x0, vae, image_encoder, image_processor, generator):
    """generate_data.py:701-705 / 743-746 -- decode (PyTorch), postprocess (identity), bicubic 224 (K8), guide (PyTorch)."""
    D_x0_t = vae.decode(pred_x0 / vae.config.scaling_factor, return_dict=False, generator=generator)[0]
    D_x0_t = image_processor.postprocess(D_x0_t, output_type="pt", do_denormalize=[False] * D_x0_t.shape[0])
    if D_x0_t.requires_grad and torch.is_grad_enabled():
        D_x0_t = ops.BicubicResize.apply(D_x0_t, (224, 224))                                      # :704 / :745
    else:
        D_x0_t = ops.bicubic_resize(D_x0_t, (224, 224))
    return image_encoder.encode_image(D_x0_t).float()


def draw_channel_noise(bs: int, channel_noise_dim: int):
    """generate_data.py:692-695 -- CPU global RNG, rand first then normal_ (host tensors; the caller moves them)."""
    channel_noise = torch.rand([bs, channel_noise_dim, 1, 1])
    channel_noise_bias = torch.zeros([bs, channel_noise_dim, 1, 1]).normal_(0, 1)
    return channel_noise, channel_noise_bias


def transform_guidance_core(latents, targets, channel_noise, channel_noise_bias, sub_timesteps, noise_scheduler, unet,
                            prompt_embeds, class_labels, vae, image_encoder, image_processor, generator,
                            total_global_proto, total_local_proto):
    """generate_data.py:696-732 given the drawn channel parameters (device fp32 leaves).  Shape-static and free of host
    synchronisation, so expand.Expander can capture it -- forward, autograd backward and the update -- in a CUDA graph."""
    a = _need_args()
    channel_noise = channel_noise.detach().requires_grad_(True)
    channel_noise_bias = channel_noise_bias.detach().requires_grad_(True)
    latents = latents.detach()
    x_dec_noisy = ops.ChannelAffine.apply(latents, channel_noise, channel_noise_bias)          # :696

    score = 0.0
    for temp_t in sub_timesteps:
        x_dec_noisy, pred_x0 = denoise_one_step(x_dec_noisy, noise_scheduler, temp_t, unet, prompt_embeds, class_labels)
        image_features = _guide_features(pred_x0, vae, image_encoder, image_processor, generator)
        score = score + ops.PrototypeEnergy.apply(image_features, targets, total_global_proto, total_local_proto,
                                                  float(a.gs), float(a.ls), False)            # :707-717
    score = score / a.guidance_period

    channel_noise_grad, channel_noise_bias_grad = torch.autograd.grad(score, [channel_noise, channel_noise_bias])
    channel_noise.data.add_(-a.rho * channel_noise_grad)                                        # :723-724
    channel_noise_bias.data.add_(-a.rho * channel_noise_bias_grad)

    # :726-728 -- transform again with the updated params and project onto the L-inf ball around the input
    latents = ops.affine_project(latents, channel_noise.data, channel_noise_bias.data, float(a.constraint_value))
    return latents.detach(), score.detach()


def transform_guidance(latents, batch, sub_timesteps, noise_scheduler, unet, prompt_embeds, class_labels,
                       vae, image_encoder, image_processor, weight_dtype, generator,
                       total_global_proto, total_local_proto):
    """generate_data.py:687-732 -> (latents, score)."""
    # :692-695 -- CPU global RNG, rand first then normal_; kept as fp32 leaves (the reference rounds them to fp16)
    channel_noise, channel_noise_bias = draw_channel_noise(latents.shape[0], latents.shape[1])
    return transform_guidance_core(latents, batch["targets"], channel_noise.to(latents.device), channel_noise_bias.to(latents.device),
                                   sub_timesteps, noise_scheduler, unet, prompt_embeds, class_labels, vae, image_encoder,
                                   image_processor, generator, total_global_proto, total_local_proto)


def direct_guidance(latents, batch, t_i, noise_scheduler, unet, prompt_embeds, class_labels,
                    vae, image_encoder, image_processor, weight_dtype, generator,
                    total_global_proto, total_local_proto):
    """generate_data.py:735-767 -> (latents, x_0, score)."""
    a = _need_args()
    cfg = bool(a.do_classifier_free_guidance)
    latents = latents.detach().requires_grad_(True)
    latent_model_input = torch.cat([latents] * 2) if cfg else latents
    noise_pred = unet(latent_model_input, t_i, prompt_embeds, class_labels=class_labels, return_dict=False)[0]
    a_t, a_prev = noise_scheduler.alpha_pair(t_i)
    # the score depends on x_0 only: run K5 for x_0 under autograd, and produce x_next - rho*grad afterwards
    # in one fused launch (K5 with its guidance epilogue) instead of materialising x_next twice
    _unused_prev, x_0 = ops.CfgDdimStep.apply(noise_pred, latents, float(a.guidance_scale), a_t, a_prev, cfg)
    image_features = _guide_features(x_0, vae, image_encoder, image_processor, generator)
    score = ops.PrototypeEnergy.apply(image_features, batch["targets"], total_global_proto, total_local_proto,
                                      float(a.gs), float(a.ls), True)                          # :747-759 (normalised f)
    x_dec_grad = torch.autograd.grad(score, latents)[0]                                         # :761
    x_dec_next, _ = ops.cfg_ddim_step(noise_pred.detach(), latents.detach(), float(a.guidance_scale), a_t, a_prev,
                                      cfg=cfg, grad=x_dec_grad, rho=float(a.rho), want_x0=False)  # :741 + :762
    return x_dec_next, x_0.detach(), score.detach()


def add_noise(noise_scheduler, model_input, noise, t_enc):
    """generate_data.py:1176."""
    return noise_scheduler.add_noise(model_input, noise, t_enc)


def start_index(strength: float, n_timesteps: int) -> int:
    """generate_data.py:1174 -- int() truncation of the float product is part of the contract."""
    return int((1 - strength) * n_timesteps)


def guide_timesteps(timesteps, guidance_step: int, guidance_period: int) -> list:
    """generate_data.py:1178-1180."""
    g = timesteps[len(timesteps) - guidance_step: len(timesteps) - guidance_step + guidance_period].tolist()
    assert len(g) == guidance_period
    assert guidance_step >= 1  # start from 1
    return g


def split_mask(total_data_number: int, split: int, total_split: int) -> list:
    """generate_data.py:1002-1007, with the documented overshoot of a non-last split clamped to the data."""
    import math
    number_per_split = math.ceil(total_data_number / total_split)
    lo = min(number_per_split * split, total_data_number)
    hi = min(number_per_split * (split + 1), total_data_number)
    return list(range(lo, hi))
