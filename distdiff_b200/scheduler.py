"""Host side of the DDIM scheduler the reference gets from diffusers (generate_data.py:863,1044,119,1176).

Only table construction and index arithmetic live here (a 1000-entry fp32 table, built once on CPU with the
same torch ops diffusers uses); the per-step tensor math is the K5 / K7 kernels.  SD-v1.4 scheduler config:
scaled_linear betas 0.00085 -> 0.012, 1000 train steps, steps_offset 1, set_alpha_to_one False,
clip_sample False, epsilon prediction, 'leading' spacing, eta 0.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class DDIMScheduler:
    """Duck-type of diffusers.DDIMScheduler restricted to what generate_data.py calls."""

    init_noise_sigma = 1.0  # generate_data.py:1158

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 steps_offset: int = 1):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)          # CPU fp32, like diffusers
        self._abar = [float(v) for v in self.alphas_cumprod]                # python floats holding the fp32 values
        self.final_alpha_cumprod = self.alphas_cumprod[0]                   # set_alpha_to_one=False
        self.num_inference_steps = None
        self.timesteps = None

    @classmethod
    def from_pretrained(cls, *_args, **_kwargs):
        """No network, no files: SD-v1.4's scheduler_config.json values are the constructor defaults."""
        return cls()

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts)   # CPU int64, like retrieve_timesteps(..., "cpu") generate_data.py:1044
        return self.timesteps

    def alpha_pair(self, t):
        """(abar_t, abar_prev) for DDIMScheduler.step at timestep t (prev_t = t - 1000 // steps)."""
        t = int(t)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        return self._abar[t], (self._abar[prev_t] if prev_t >= 0 else self._abar[0])

    def scale_model_input(self, sample, timestep=None):
        return sample  # identity for DDIM (generate_data.py:111)

    def step(self, model_output, timestep, sample, return_dict=True):
        """diffusers signature; one K5 launch (no CFG: model_output is already the combined epsilon)."""
        a_t, a_prev = self.alpha_pair(timestep)
        prev, x0 = ops.cfg_ddim_step(model_output, sample, 1.0, a_t, a_prev, cfg=False)
        if return_dict:
            return {"prev_sample": prev, "pred_original_sample": x0}
        return (prev,)

    def add_noise(self, original_samples, noise, timesteps):
        """One K7 launch.  generate_data.py:1176 passes a single timestep for the whole batch."""
        t = int(timesteps.reshape(-1)[0]) if isinstance(timesteps, torch.Tensor) else int(timesteps)
        return ops.add_noise(original_samples, noise, self._abar[t])


def retrieve_timesteps(scheduler: DDIMScheduler, num_inference_steps: int, device=None):
    """diffusers' helper used at generate_data.py:1044."""
    return scheduler.set_timesteps(num_inference_steps, device=device), num_inference_steps
