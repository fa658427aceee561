"""BENCHMARK COMPARATOR ONLY -- the PyTorch-eager op sequences the sm_100a kernels replace, on the same GPU.

SURVEY.md section 2c/6 sets the bar for the hand-written kernels as "beat the eager sequence on the same B200".  This
module restates those sequences literally (generate_data.py:110-120 + diffusers' DDIMScheduler.step, :696, :124-137,
:707-717 / :747-759, :704, :1176, :1227-1234) behind the names ``distdiff_b200.ops`` exports, so that
``bench.py --ops eager`` and the ``eager_*`` rows of the micro-benchmark can run the SAME host code with the fused
kernels swapped for ATen launches.  It is never imported by the product path (guidance / expand / generate_data use
``ops``; tests/test_cabi_loads.py checks that), and it is not a fallback: nothing selects it automatically.

Like diffusers, the scheduler scalars are 0-dim fp32 device tensors (alphas_cumprod lives on the device there), so the
step issues the same ~13 small launches the reference does.
"""
from __future__ import annotations

import contextlib

import torch

launch_count = 0   # not tracked for ATen launches (bench.py reports the profiler's kernel count instead)
profiler = None


_scalars = {}


def _dev_scalar(v, ref):
    """0-dim fp32 device tensor holding ``v`` -- diffusers indexes a device-resident alphas_cumprod table, which costs no
    host-to-device copy; cached so that the same holds here (and under CUDA-graph capture, after the warm-up runs)."""
    key = (float(v), ref.device)
    t = _scalars.get(key)
    if t is None:
        t = _scalars[key] = torch.tensor(float(v), dtype=torch.float32, device=ref.device)
    return t


# ---- K5: CFG combine + DDIMScheduler.step (eta = 0, epsilon prediction) ------------------------------------------
def cfg_ddim_step(noise_pred, x, guidance_scale, a_t, a_prev, cfg=True, grad=None, rho=0.0, want_prev=True, want_x0=True):
    if cfg:
        noise_pred_uncond, noise_pred_text = noise_pred.chunk(2)                                    # :115
        noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)    # :116
    alpha_prod_t, alpha_prod_t_prev = _dev_scalar(a_t, x), _dev_scalar(a_prev, x)                   # alphas_cumprod[t] on device
    beta_prod_t = 1 - alpha_prod_t
    pred_original_sample = (x - beta_prod_t ** 0.5 * noise_pred) / alpha_prod_t ** 0.5
    pred_sample_direction = (1 - alpha_prod_t_prev) ** 0.5 * noise_pred
    prev_sample = alpha_prod_t_prev ** 0.5 * pred_original_sample + pred_sample_direction
    prev_sample, pred_original_sample = prev_sample.to(x.dtype), pred_original_sample.to(x.dtype)
    if grad is not None:
        prev_sample = prev_sample - rho * grad                                                      # :762
    return (prev_sample if want_prev else None), (pred_original_sample if want_x0 else None)


class CfgDdimStep:
    @staticmethod
    def apply(noise_pred, x, guidance_scale, a_t, a_prev, cfg):
        return cfg_ddim_step(noise_pred, x, guidance_scale, a_t, a_prev, cfg)


# ---- K6: channel affine (+ L-inf projection, generate_data.py:124-137) --------------------------------------------
def tensor_clamp(t, min, max, in_place=True):
    res = t if in_place else t.clone()
    idx = res.data < min
    res.data[idx] = min[idx]
    idx = res.data > max
    res.data[idx] = max[idx]
    return res


def affine_project(x, a, b, radius=-1.0, center=None):
    y = x * (1 + a.to(x.dtype)) + b.to(x.dtype)                                                     # :696 / :727
    if radius >= 0:
        c = x.clone() if center is None else center                                                 # :726
        tensor_clamp(y, min=c - radius, max=c + radius, in_place=True)                              # :728
    return y


class ChannelAffine:
    @staticmethod
    def apply(x, a, b):
        return x * (1 + a.to(x.dtype)) + b.to(x.dtype)


# ---- K4: prototype energy (autograd does the backward) ------------------------------------------------------------
class PrototypeEnergy:
    @staticmethod
    def apply(image_features, targets, total_global_proto, total_local_proto, gs, ls, normalize_f):
        score = 0.0
        if normalize_f:
            image_features = image_features / image_features.norm(dim=-1, keepdim=True)            # :747
        if total_global_proto is not None:
            global_proto = total_global_proto[targets]
            global_distance = torch.norm(image_features - global_proto.detach(), dim=1, p=2).mean()
            score += global_distance * gs
        if total_local_proto is not None:
            local_proto = total_local_proto[targets]
            target_cluster_index = torch.argmax(torch.bmm(image_features.unsqueeze(1), local_proto.permute(0, 2, 1)), -1)
            local_proto = local_proto[torch.arange(local_proto.size(0)), target_cluster_index.reshape(-1)]
            local_distance = torch.norm(image_features - local_proto.detach(), dim=1, p=2).mean()
            score += local_distance * ls
        return score


def energy_fwd_bwd(f, targets, g, l, gs, ls, normalize_f=False, mode="auto"):
    """forward + autograd backward of the literal sequence -> (score, None, None, grad) like ops.energy_fwd_bwd"""
    fr = f.detach().requires_grad_(True)
    s = PrototypeEnergy.apply(fr, targets, g, l, gs, ls, normalize_f)
    (grad,) = torch.autograd.grad(s, fr)
    return s.detach(), None, None, grad


# ---- K8 / K9 / K7 -------------------------------------------------------------------------------------------------
def bicubic_resize(x, size):
    return torch.nn.functional.interpolate(x, size=size, mode="bicubic")                            # :704


class BicubicResize:
    @staticmethod
    def apply(x, size):
        return bicubic_resize(x, size)


def image_to_uint8(image, denormalize=True, out=None):
    x = (image / 2 + 0.5).clamp(0, 1) if denormalize else image                                     # :1227
    return x.mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8).contiguous()     # save_image


def add_noise(x, noise, a_t):
    alpha = _dev_scalar(a_t, x)
    sqrt_alpha_prod = (alpha ** 0.5).to(x.dtype)
    sqrt_one_minus_alpha_prod = ((1 - alpha) ** 0.5).to(x.dtype)
    return sqrt_alpha_prod * x + sqrt_one_minus_alpha_prod * noise                                  # diffusers add_noise


# ---- prototype-stage sequences (dataloader.py:677-707 on the GPU; a k-means iteration as cdist/argmin/index_add_) ----
def rownorm_classsum(feat, labels, num_classes):
    fn = feat / feat.norm(dim=-1, keepdim=True)                                                      # :677
    csum = torch.zeros(num_classes, feat.shape[1], dtype=torch.float32, device=feat.device).index_add_(0, labels, fn)
    cnt = torch.bincount(labels, minlength=num_classes)
    return fn, csum, cnt


def kmeans_iteration(x_sorted, row_class, centroid):
    """x_sorted [N,D], row_class [N] (class of each row), centroid [C,K,D]: assignments + new sums, eager."""
    C, K, D = centroid.shape
    cen = centroid[row_class]                                                                        # [N,K,D] gather
    d = torch.cdist(x_sorted.unsqueeze(1), cen).squeeze(1)                                           # [N,K]
    a = d.argmin(-1)
    flat = row_class * K + a
    sums = torch.zeros(C * K, D, dtype=torch.float32, device=x_sorted.device).index_add_(0, flat, x_sorted)
    cnt = torch.bincount(flat, minlength=C * K)
    return a, sums.view(C, K, D), cnt.view(C, K)


@contextlib.contextmanager
def swapped():
    """Run the host code (guidance / expand / scheduler) with the fused kernels replaced by the eager sequences."""
    import sys
    from . import expand, guidance, scheduler
    me = sys.modules[__name__]
    saved = (guidance.ops, expand.ops, scheduler.ops)
    guidance.ops = expand.ops = scheduler.ops = me
    try:
        yield me
    finally:
        guidance.ops, expand.ops, scheduler.ops = saved
