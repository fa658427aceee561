"""Oracle: hierarchical-prototype energy (TEST INFRASTRUCTURE ONLY).

``energy_score`` is a literal restatement of generate_data.py:707-717 (transform
mode: features NOT normalised) and :747-759 (direct mode: features L2-normalised
at :747), in torch CPU ops so ``torch.autograd`` gives the reference gradient.

``energy_fwd_bwd`` is the closed form the CUDA kernel implements:
    score = gs * mean_b ||f_b - g_y|| + ls * mean_b ||f_b - l_{y,k*}||,
    k*    = argmax_k <f_b, l_{y,k}>            (first max on ties, like torch.argmax)
    dscore/df_b = (gs/B) (f_b - g_y)/||.|| + (ls/B) (f_b - l*)/||.||   (0 where the norm is 0,
                                                 torch.norm's sub-gradient)
and, when ``normalize_f`` (direct mode), chained through f -> f/||f||.
"""
from __future__ import annotations

import numpy as np
import torch


def energy_score(image_features, targets, total_global_proto, total_local_proto, gs, ls, normalize_f=False):
    """Literal: generate_data.py:705-717 / 746-759.  ``targets`` is a python list like batch["targets"]."""
    score = 0.0
    if normalize_f:
        image_features = image_features / image_features.norm(dim=-1, keepdim=True)  # :747
    if total_global_proto is not None:
        global_proto = total_global_proto[targets]
        global_distance = torch.norm(image_features - global_proto.detach(), dim=1, p=2).mean()
        score += global_distance * gs
    if total_local_proto is not None:
        local_proto = total_local_proto[targets]
        target_cluster_index = torch.argmax(
            torch.bmm(image_features.unsqueeze(1), local_proto.permute(0, 2, 1)), -1)
        # the reference's .squeeze() breaks for B == 1 (0-d index -> [1,K,D] selection still broadcasts
        # correctly in the subtraction); reshape(-1) is the same thing for B >= 2 and safe for B == 1.
        local_proto = local_proto[torch.arange(local_proto.size(0)), target_cluster_index.reshape(-1)]
        local_distance = torch.norm(image_features - local_proto.detach(), dim=1, p=2).mean()
        score += local_distance * ls
    return score


def energy_fwd_bwd(f, targets, g, l, gs, ls, normalize_f=False, dtype=np.float64):
    """Closed-form score, per-sample distances, k* and d score / d f (numpy, ``dtype`` arithmetic)."""
    f = np.asarray(f, dtype=dtype)
    B, D = f.shape
    y = np.asarray(targets, dtype=np.int64)
    fn = f
    if normalize_f:
        nrm = np.sqrt((f * f).sum(-1, keepdims=True))
        fn = f / nrm
    grad_fn = np.zeros_like(fn)
    per = np.zeros((B, 2), dtype=dtype)
    kstar = np.zeros(B, dtype=np.int32)
    score = dtype(0)
    if g is not None:
        diff = fn - np.asarray(g, dtype=dtype)[y]
        d = np.sqrt((diff * diff).sum(-1))
        per[:, 0] = d
        score += gs * d.mean()
        safe = np.where(d > 0, d, 1)
        grad_fn += (gs / B) * np.where(d[:, None] > 0, diff / safe[:, None], 0)
    if l is not None:
        ly = np.asarray(l, dtype=dtype)[y]  # [B,K,D]
        dots = np.einsum("bd,bkd->bk", fn, ly)
        kstar = dots.argmax(-1).astype(np.int32)  # first max on ties
        diff = fn - ly[np.arange(B), kstar]
        d = np.sqrt((diff * diff).sum(-1))
        per[:, 1] = d
        score += ls * d.mean()
        safe = np.where(d > 0, d, 1)
        grad_fn += (ls / B) * np.where(d[:, None] > 0, diff / safe[:, None], 0)
    if normalize_f:
        grad_f = (grad_fn - fn * (fn * grad_fn).sum(-1, keepdims=True)) / nrm
    else:
        grad_f = grad_fn
    return score, per, kstar, grad_f
