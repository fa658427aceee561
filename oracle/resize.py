"""Oracle: the decoded-image resize in front of the guide network (TEST INFRASTRUCTURE ONLY).

generate_data.py:704 / :745 call ``torch.nn.functional.interpolate(D_x0_t, size=(224, 224), mode='bicubic')``;
``interpolate_bicubic`` below IS that call on the CPU (the reference's own third-party function).
``bicubic_numpy`` restates ATen's upsample_bicubic2d from scratch (what the CUDA kernel implements): source
coordinate scale*(dst+0.5)-0.5 with scale = in/out, cubic convolution with A = -0.75, taps clamped to the image,
x first then y.  tests/test_oracle_golden.py pins the restatement against the torch call.
"""
from __future__ import annotations

import numpy as np


def interpolate_bicubic(x, size):
    import torch
    return torch.nn.functional.interpolate(x, size=size, mode="bicubic")


def _coeffs(t):
    A = np.float32(-0.75)
    t = np.float32(t)
    x1 = t
    x1p = np.float32(x1 + np.float32(1))
    c0 = ((A * x1p - np.float32(5) * A) * x1p + np.float32(8) * A) * x1p - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x1 - (A + np.float32(3))) * x1 * x1 + np.float32(1)
    x2 = np.float32(np.float32(1) - t)
    x2p = np.float32(x2 + np.float32(1))
    c2 = ((A + np.float32(2)) * x2 - (A + np.float32(3))) * x2 * x2 + np.float32(1)
    c3 = ((A * x2p - np.float32(5) * A) * x2p + np.float32(8) * A) * x2p - np.float32(4) * A
    return np.array([c0, c1, c2, c3], dtype=np.float32)


def axis_matrix(n_in: int, n_out: int) -> np.ndarray:
    """[n_out, n_in] fp32 interpolation matrix of one axis (rows sum to 1)."""
    scale = np.float32(n_in) / np.float32(n_out)
    M = np.zeros((n_out, n_in), dtype=np.float32)
    for o in range(n_out):
        real = np.float32(scale * np.float32(o + 0.5) - np.float32(0.5))
        f = int(np.floor(real))
        c = _coeffs(real - np.float32(f))
        for k in range(4):
            M[o, min(max(f - 1 + k, 0), n_in - 1)] += c[k]
    return M


def bicubic_numpy(x: np.ndarray, size) -> np.ndarray:
    """x [B,C,H,W] -> [B,C,Ho,Wo] through the two axis matrices (fp64 accumulation of fp32 weights)."""
    ho, wo = size
    My = axis_matrix(x.shape[2], ho).astype(np.float64)
    Mx = axis_matrix(x.shape[3], wo).astype(np.float64)
    return np.matmul(np.matmul(My, x.astype(np.float64)), Mx.T)                     # [Ho,H] @ [B,C,H,W] @ [W,Wo]


def bicubic_backward_numpy(g: np.ndarray, in_hw) -> np.ndarray:
    """Transpose of the linear map above: grad wrt the input."""
    H, W = in_hw
    My = axis_matrix(H, g.shape[2]).astype(np.float64)
    Mx = axis_matrix(W, g.shape[3]).astype(np.float64)
    return np.matmul(np.matmul(My.T, g.astype(np.float64)), Mx)                    # [H,Ho] @ [B,C,Ho,Wo] @ [Wo,W]
