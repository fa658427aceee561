"""Oracle: scheduler arithmetic of the sampling step (TEST INFRASTRUCTURE ONLY).

Restates, on CPU in fp32 torch ops:

* ``denoise_one_step``'s CFG combine + DDIM step  -- generate_data.py:109-121
* the SD-v1.4 ``DDIMScheduler`` contract used at generate_data.py:863,1044,119,1176
  (diffusers is a pip dependency absent from /root/reference and not pinned:
  INSTALL.md:31-33 installs git main, floor ``check_min_version("0.28.0.dev0")``
  generate_data.py:74).  Published algorithm restated here:
    betas      = linspace(sqrt(0.00085), sqrt(0.012), 1000, f32) ** 2   ("scaled_linear")
    abar       = cumprod(1 - betas)
    timesteps  = (arange(50) * (1000 // 50))[::-1] + steps_offset(1)     ("leading")
    prev_t     = t - 1000 // 50 ; abar_prev = abar[prev_t] if prev_t >= 0 else abar[0]
                 (set_alpha_to_one=False)
    x0         = (x - sqrt(1 - abar_t) * eps) / sqrt(abar_t)             (epsilon pred., no clip)
    x_prev     = sqrt(abar_prev) * x0 + sqrt(1 - abar_prev) * eps        (eta = 0)
    add_noise  = sqrt(abar_t) * x + sqrt(1 - abar_t) * noise
* timestep-index arithmetic of main()             -- generate_data.py:1174-1180
* ``--split/--total_split`` sharding              -- generate_data.py:1002-1009
"""
from __future__ import annotations

import math

import numpy as np
import torch

NUM_TRAIN_TIMESTEPS = 1000
BETA_START = 0.00085
BETA_END = 0.012
STEPS_OFFSET = 1


def alphas_cumprod() -> torch.Tensor:
    """SD-v1.4 scheduler_config.json: scaled_linear betas, 1000 train steps (fp32)."""
    betas = torch.linspace(BETA_START ** 0.5, BETA_END ** 0.5, NUM_TRAIN_TIMESTEPS, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def timesteps(num_inference_steps: int = 50) -> torch.Tensor:
    """``retrieve_timesteps(noise_scheduler, 50, 'cpu')`` generate_data.py:1043-1044 -> [981, ..., 1]."""
    step_ratio = NUM_TRAIN_TIMESTEPS // num_inference_steps
    ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
    return torch.from_numpy(ts + STEPS_OFFSET)


def alpha_pair(t: int, num_inference_steps: int = 50, abar: torch.Tensor | None = None):
    """(abar_t, abar_prev) as used by DDIMScheduler.step for timestep ``t``."""
    abar = alphas_cumprod() if abar is None else abar
    prev_t = int(t) - NUM_TRAIN_TIMESTEPS // num_inference_steps
    a_t = abar[int(t)]
    a_prev = abar[prev_t] if prev_t >= 0 else abar[0]  # set_alpha_to_one=False -> final_alpha_cumprod = abar[0]
    return a_t, a_prev


def cfg_combine(noise_pred_2b: torch.Tensor, guidance_scale: float) -> torch.Tensor:
    """generate_data.py:115-117."""
    noise_pred_uncond, noise_pred_text = noise_pred_2b.chunk(2)
    return noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)


def ddim_step(model_output: torch.Tensor, sample: torch.Tensor, a_t, a_prev):
    """diffusers DDIMScheduler.step with eta=0, epsilon prediction, no clipping/thresholding.

    Returns (prev_sample, pred_original_sample) -- generate_data.py:119-120.
    """
    beta_prod_t = 1 - a_t
    pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / a_t ** 0.5
    pred_sample_direction = (1 - a_prev) ** 0.5 * model_output
    prev_sample = a_prev ** 0.5 * pred_original_sample + pred_sample_direction
    return prev_sample, pred_original_sample


def cfg_ddim_step(noise_pred_2b, latents, guidance_scale, a_t, a_prev, grad=None, rho=0.0):
    """CFG combine + DDIM step (+ optional ``x_prev - rho * grad``, generate_data.py:762)."""
    eps = cfg_combine(noise_pred_2b, guidance_scale)
    prev, x0 = ddim_step(eps, latents, a_t, a_prev)
    if grad is not None:
        prev = prev - rho * grad
    return prev, x0


def add_noise(original, noise, a_t):
    """diffusers DDIMScheduler.add_noise -- generate_data.py:1176."""
    return a_t ** 0.5 * original + (1 - a_t) ** 0.5 * noise


def start_index(strength: float, n_timesteps: int = 50) -> int:
    """generate_data.py:1174 -- truncation of a float product (0.9 -> 4, 0.8 -> 9, 0.5 -> 25)."""
    return int((1 - strength) * n_timesteps)


def guide_timesteps(ts: torch.Tensor, guidance_step: int, guidance_period: int) -> list:
    """generate_data.py:1178-1180."""
    g = ts[len(ts) - guidance_step: len(ts) - guidance_step + guidance_period].tolist()
    assert len(g) == guidance_period
    assert guidance_step >= 1
    return g


def split_mask(total_data_number: int, split: int, total_split: int) -> list:
    """generate_data.py:1002-1007 (literal; a non-last split may run past the end -- Subset would raise)."""
    number_per_split = math.ceil(total_data_number / total_split)
    if split == (total_split - 1) and total_data_number < number_per_split * (split + 1):
        return list(range(number_per_split * split, total_data_number))
    return list(range(number_per_split * split, number_per_split * (split + 1)))


class OracleDDIMScheduler:
    """Duck-typed stand-in for diffusers.DDIMScheduler (only what generate_data.py calls)."""

    init_noise_sigma = 1.0

    def __init__(self, num_inference_steps: int = 50):
        self.alphas_cumprod = alphas_cumprod()
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_inference_steps = num_inference_steps
        self.timesteps = timesteps(num_inference_steps)

    def scale_model_input(self, sample, timestep=None):  # generate_data.py:111 -- identity for DDIM
        return sample

    def step(self, model_output, timestep, sample, return_dict=True):
        a_t, a_prev = alpha_pair(int(timestep), self.num_inference_steps, self.alphas_cumprod)
        a_t = a_t.to(sample.device)
        a_prev = a_prev.to(sample.device)
        prev, x0 = ddim_step(model_output, sample, a_t, a_prev)
        return {"prev_sample": prev, "pred_original_sample": x0}

    def add_noise(self, original_samples, noise, timesteps):
        a_t = self.alphas_cumprod.to(dtype=original_samples.dtype)[int(timesteps)]
        return add_noise(original_samples, noise, a_t)
