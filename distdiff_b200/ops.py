"""Torch-tensor front end of the C ABI (include/distdiff_sm100.h).

PyTorch is used for device memory, streams and autograd bookkeeping only; every arithmetic op on the
hot path is a kernel of libdistdiff_sm100.so.  No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import DistDiffError, check

_DTYPES = {torch.float32: _lib.DD_F32, torch.float16: _lib.DD_F16, torch.bfloat16: _lib.DD_BF16}

# launches of OUR kernels since process start (bench.py reports the delta over the timed region)
launch_count = 0


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


# optional per-launch timing: set to a list to collect (entry point, algorithmic bytes, start event, end event)
# for every call issued outside CUDA-graph capture (bench.py's roofline leg); None = no events, no overhead
profiler = None


def _call(name: str, algo_bytes: int, n_kernels: int, *args) -> None:
    """Invoke one C-ABI entry point on the current stream, raise on a non-zero return."""
    fn = getattr(_lib.lib(), name)
    if profiler is not None and not torch.cuda.is_current_stream_capturing():
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        profiler.append((name, int(algo_bytes), e0, e1))
    else:
        rc = fn(*args)
    check(rc, name)
    _count(n_kernels)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise DistDiffError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise DistDiffError(f"{name}: CUDA tensor required (the guidance hot path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise DistDiffError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise DistDiffError(f"unsupported latent dtype {t.dtype} (f32 / f16 / bf16)") from None


# ------------------------------------------------------------------------------------------------
# K5  CFG + DDIM step
# ------------------------------------------------------------------------------------------------
def _split_noise(noise_pred: torch.Tensor, x: torch.Tensor, cfg: bool):
    noise_pred = _req(noise_pred, "noise_pred", x.dtype)
    n = x.numel()
    if cfg:
        if noise_pred.numel() != 2 * n:
            raise DistDiffError(f"noise_pred has {noise_pred.numel()} elements, expected 2 x {n} (CFG batch)")
        es = noise_pred.element_size()
        return noise_pred, C.c_void_p(noise_pred.data_ptr()), C.c_void_p(noise_pred.data_ptr() + n * es)
    if noise_pred.numel() != n:
        raise DistDiffError(f"noise_pred has {noise_pred.numel()} elements, expected {n}")
    return noise_pred, C.c_void_p(noise_pred.data_ptr()), None


def cfg_ddim_step(noise_pred: torch.Tensor, x: torch.Tensor, guidance_scale: float, a_t: float, a_prev: float,
                  cfg: bool = True, grad: Optional[torch.Tensor] = None, rho: float = 0.0,
                  want_prev: bool = True, want_x0: bool = True):
    """No-grad fused step.  noise_pred: [2B,...] (uncond first, text second) when ``cfg`` else [B,...].
    Returns (x_prev, x0); entries not wanted are None.  generate_data.py:115-120 (+ :762 with grad/rho)."""
    x = _req(x, "latents")
    keep, pu, pt = _split_noise(noise_pred, x, cfg)
    if grad is not None:
        grad = _req(grad, "grad", x.dtype)
    x_prev = torch.empty_like(x) if want_prev else None
    x0 = torch.empty_like(x) if want_x0 else None
    _call("dd_cfg_ddim_fwd", x.numel() * x.element_size() * (2 + int(pt is not None) + int(grad is not None) + int(want_prev) + int(want_x0)), 1,
          pu, pt, _ptr(x), x.numel(), _code(x), float(guidance_scale), float(a_t),
                                     float(a_prev), _ptr(grad), float(rho), _ptr(x_prev), _ptr(x0), _stream())
    del keep
    return x_prev, x0


class CfgDdimStep(torch.autograd.Function):
    """Differentiable K5: (noise_pred[2B], x) -> (x_prev, x0).  Backward is one dd_cfg_ddim_bwd launch."""

    @staticmethod
    def forward(ctx, noise_pred, x, guidance_scale, a_t, a_prev, cfg):
        x_prev, x0 = cfg_ddim_step(noise_pred.detach(), x.detach(), guidance_scale, a_t, a_prev, cfg)
        ctx.consts = (float(guidance_scale), float(a_t), float(a_prev), bool(cfg))
        ctx.meta = (noise_pred.shape, x.shape, x.dtype)
        return x_prev, x0

    @staticmethod
    def backward(ctx, g_prev, g_x0):
        s, a_t, a_prev, cfg = ctx.consts
        np_shape, x_shape, dtype = ctx.meta
        need_np, need_x = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if g_prev is None and g_x0 is None:
            return None, None, None, None, None, None
        ref = g_prev if g_prev is not None else g_x0
        g_prev = None if g_prev is None else _req(g_prev, "g_prev", dtype)
        g_x0 = None if g_x0 is None else _req(g_x0, "g_x0", dtype)
        n = ref.numel()
        g_np = torch.empty(np_shape, dtype=dtype, device=ref.device) if need_np else None
        g_x = torch.empty(x_shape, dtype=dtype, device=ref.device) if need_x else None
        pu = pt = None
        if g_np is not None:
            pu = C.c_void_p(g_np.data_ptr())
            pt = C.c_void_p(g_np.data_ptr() + n * g_np.element_size()) if cfg else None
        _call("dd_cfg_ddim_bwd", n * ref.element_size() * (int(g_prev is not None) + int(g_x0 is not None) + (2 if cfg else 1) * int(g_np is not None) + int(g_x is not None)), 1,
              _ptr(g_prev), _ptr(g_x0), n, _DTYPES[dtype], s, a_t, a_prev, int(cfg), pu, pt,
                                         _ptr(g_x), _stream())
        return g_np, g_x, None, None, None, None


# ------------------------------------------------------------------------------------------------
# K6  channel affine + L-inf projection
# ------------------------------------------------------------------------------------------------
def affine_project(x: torch.Tensor, a: torch.Tensor, b: torch.Tensor, radius: float = -1.0,
                   center: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x*(1+a)+b per (batch, channel); clamp to center +- radius when radius >= 0 (center defaults to x).
    generate_data.py:696, :726-728, :124-137.  a, b: [B,C,1,1] (any float dtype; fp32 inside)."""
    x = _req(x, "latents")
    if x.dim() < 2:
        raise DistDiffError("latents must be [B, C, ...]")
    BC = x.shape[0] * x.shape[1]
    HW = x.numel() // max(BC, 1)
    a32 = _req(a, "channel_noise").reshape(-1).float().contiguous()
    b32 = _req(b, "channel_noise_bias").reshape(-1).float().contiguous()
    if a32.numel() != BC or b32.numel() != BC:
        raise DistDiffError(f"channel params must have B*C = {BC} elements")
    if center is not None:
        center = _req(center, "center", x.dtype)
        if center.shape != x.shape:
            raise DistDiffError("center must have the shape of the latents")
    y = torch.empty_like(x)
    _call("dd_affine_project_fwd", x.numel() * x.element_size() * (2 + int(center is not None and radius >= 0)), 1,
          _ptr(x), _ptr(a32), _ptr(b32), _ptr(center), BC, HW, _code(x), float(radius),
                                           _ptr(y), _stream())
    return y


class ChannelAffine(torch.autograd.Function):
    """Differentiable y = x*(1+a)+b (no clamp) -- the graph root of transform_guidance (generate_data.py:696)."""

    @staticmethod
    def forward(ctx, x, a, b):
        y = affine_project(x.detach(), a.detach(), b.detach(), -1.0)
        ctx.save_for_backward(x, a)
        ctx.pshape = (a.shape, a.dtype, b.shape, b.dtype)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, a = ctx.saved_tensors
        x = _req(x, "latents")
        g_y = _req(g_y, "g_y", x.dtype)
        BC = x.shape[0] * x.shape[1]
        HW = x.numel() // BC
        a32 = a.detach().reshape(-1).float().contiguous()
        g_a = torch.empty(BC, dtype=torch.float32, device=x.device)
        g_b = torch.empty(BC, dtype=torch.float32, device=x.device)
        g_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        _call("dd_affine_bwd", x.numel() * x.element_size() * (2 + int(g_x is not None)), 1,
              _ptr(g_y), _ptr(x), _ptr(a32), BC, HW, _code(x), _ptr(g_a), _ptr(g_b), _ptr(g_x),
                                       _stream())
        ash, adt, bsh, bdt = ctx.pshape
        return g_x, g_a.reshape(ash).to(adt), g_b.reshape(bsh).to(bdt)


# ------------------------------------------------------------------------------------------------
# K8  bicubic resize (decoded image -> guide input)
# ------------------------------------------------------------------------------------------------
def _hw(size):
    if isinstance(size, int):
        return size, size
    h, w = size
    return int(h), int(w)


def bicubic_resize(x: torch.Tensor, size) -> torch.Tensor:
    """torch.nn.functional.interpolate(x, size=size, mode='bicubic') for [B,C,H,W] (generate_data.py:704,745)."""
    x = _req(x, "image")
    if x.dim() != 4:
        raise DistDiffError(f"bicubic_resize expects [B,C,H,W], got {tuple(x.shape)}")
    ho, wo = _hw(size)
    B, Cc, H, W = x.shape
    out = torch.empty((B, Cc, ho, wo), dtype=x.dtype, device=x.device)
    _call("dd_bicubic_resize_fwd", (x.numel() + out.numel()) * x.element_size(), 1,
          _ptr(x), B * Cc, H, W, ho, wo, _code(x), _ptr(out), _stream())
    return out


def bicubic_resize_bwd(grad_out: torch.Tensor, in_hw) -> torch.Tensor:
    """Gradient of bicubic_resize w.r.t. its input (deterministic gather)."""
    g = _req(grad_out, "grad_out")
    B, Cc, ho, wo = g.shape
    H, W = _hw(in_hw)
    gin = torch.empty((B, Cc, H, W), dtype=g.dtype, device=g.device)
    _call("dd_bicubic_resize_bwd", (g.numel() + gin.numel()) * g.element_size(), 1,
          _ptr(g), B * Cc, H, W, ho, wo, _code(g), _ptr(gin), _stream())
    return gin


class BicubicResize(torch.autograd.Function):
    """Differentiable K8; the resize is linear, so nothing is saved for the backward."""

    @staticmethod
    def forward(ctx, x, size):
        ctx.in_hw = (x.shape[2], x.shape[3])
        return bicubic_resize(x.detach(), size)

    @staticmethod
    def backward(ctx, g):
        return bicubic_resize_bwd(g, ctx.in_hw), None


# ------------------------------------------------------------------------------------------------
# K9  decoded image -> uint8 HWC (what the PNG encoder reads)
# ------------------------------------------------------------------------------------------------
def image_to_uint8(image: torch.Tensor, denormalize: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B,C,H,W] decoder output -> [B,H,W,C] uint8, bit-identical to postprocess(do_denormalize) + save_image's
    mul(255).add_(0.5).clamp_(0,255).permute(1,2,0).to(uint8) in the image's dtype (generate_data.py:1227,1234)."""
    x = _req(image, "image")
    if x.dim() != 4:
        raise DistDiffError(f"image_to_uint8 expects [B,C,H,W], got {tuple(x.shape)}")
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=torch.uint8, device=x.device)
    elif out.shape != (B, H, W, Cc) or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
        raise DistDiffError("image_to_uint8: out must be a contiguous CUDA uint8 [B,H,W,C] tensor")
    _call("dd_image_to_uint8", x.numel() * (x.element_size() + 1), 1,
          _ptr(x), B, Cc, H, W, _code(x), int(bool(denormalize)), _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# K7  add_noise
# ------------------------------------------------------------------------------------------------
def add_noise(x: torch.Tensor, noise: torch.Tensor, a_t: float) -> torch.Tensor:
    """sqrt(abar_t) x + sqrt(1 - abar_t) noise  (diffusers add_noise at generate_data.py:1176)."""
    x = _req(x, "original_samples")
    noise = _req(noise, "noise", x.dtype)
    if noise.shape != x.shape:
        raise DistDiffError("noise must have the shape of the samples")
    out = torch.empty_like(x)
    _call("dd_add_noise", 3 * x.numel() * x.element_size(), 1,
          _ptr(x), _ptr(noise), x.numel(), _code(x), float(a_t), _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# K4  prototype energy
# ------------------------------------------------------------------------------------------------
_energy_ws = {}


def _energy_workspace(device: torch.device, B: int, C_: int) -> torch.Tensor:
    """Zero-initialised K4 workspace (ticket + class-bucketing scratch), one per (device, stream), grown on demand.
    Old buffers stay referenced: a captured CUDA graph may still hold their address."""
    key = (device.index, _stream())
    need = int(_lib.lib().dd_energy_workspace_bytes(B, C_))
    held = _energy_ws.setdefault(key, [])
    if not held or held[-1].numel() < need:
        held.append(torch.zeros(max(need, 4096), dtype=torch.uint8, device=device))
    return held[-1]


def targets_tensor(targets, C_: int, device) -> torch.Tensor:
    """batch["targets"] (a python list in the reference, generate_data.py:650) -> validated device int64."""
    if isinstance(targets, torch.Tensor):
        t = targets.to(device=device, dtype=torch.int64)
        return t if t.is_contiguous() else t.contiguous()
    ts = [int(v) for v in targets]
    if any(v < 0 or v >= C_ for v in ts):
        raise DistDiffError(f"target class out of range [0, {C_})")
    return torch.tensor(ts, dtype=torch.int64, device=device)


ENERGY_MODES = {"auto": 0, "sample": 1, "tile": 2, "tile_cta": 3, "tile_pair": 4}


def energy_fwd_bwd(f: torch.Tensor, targets, g: Optional[torch.Tensor], l: Optional[torch.Tensor], gs: float, ls: float,
                   normalize_f: bool = False, mode: str = "auto"):
    """(score[0-dim f32], per_sample [B,2], kstar [B] i32, d score/d f [B,D]) -- forward AND analytic gradient.
    generate_data.py:707-717 (normalize_f=False) / :747-759 (True).  ``mode``: "auto" (by size), "sample" (one CTA per
    sample, one launch), "tile" (class-bucketed persistent kernels for large batches; picks between the next two),
    "tile_cta" (thread-group kernel, tables in registers: any K / missing table, best below ~64 samples per SM) or
    "tile_pair" (warp-pair kernel, tables in shared memory: K <= 10 and both tables, the large-batch kernel)."""
    f = _req(f, "image_features", torch.float32)
    if f.dim() != 2:
        raise DistDiffError("image_features must be [B, D]")
    B, D = f.shape
    if g is None and l is None:
        raise DistDiffError("at least one of the global / local prototype tables is required")
    Cn = g.shape[0] if g is not None else l.shape[0]
    K = 1
    if g is not None:
        g = _req(g, "global prototypes", torch.float32)
        if g.shape != (Cn, D):
            raise DistDiffError(f"global prototypes must be [C, {D}]")
    if l is not None:
        l = _req(l, "local prototypes", torch.float32)
        if l.dim() != 3 or l.shape[0] != Cn or l.shape[2] != D:
            raise DistDiffError(f"local prototypes must be [C, K, {D}]")
        K = l.shape[1]
    y = targets_tensor(targets, Cn, f.device)
    if y.numel() != B:
        raise DistDiffError("one target per sample required")
    score = torch.empty(1, dtype=torch.float32, device=f.device)
    per = torch.empty(B, 2, dtype=torch.float32, device=f.device)
    kstar = torch.empty(B, dtype=torch.int32, device=f.device)
    grad = torch.empty_like(f)
    ws = _energy_workspace(f.device, B, Cn)
    # algorithmic bytes (compulsory HBM traffic): f read + grad written + the prototype tables once
    tiled = mode in ("tile", "tile_cta", "tile_pair") or (mode == "auto" and D <= 2048 and
                               B >= (7 if K >= 5 else 16) * torch.cuda.get_device_properties(f.device).multi_processor_count)
    _call("dd_energy_fwd_bwd", 2 * B * D * 4 + (int(g is not None) + (K if l is not None else 0)) * Cn * D * 4, 3 if tiled else 1,
          _ptr(f), _ptr(y), _ptr(g), _ptr(l), B, D, Cn, K, float(gs), float(ls),
                                       int(bool(normalize_f)), _ptr(score), _ptr(per), _ptr(kstar), _ptr(grad),
                                       _ptr(ws), ws.numel(), ENERGY_MODES[mode], _stream())
    return score.reshape(()), per, kstar, grad


class PrototypeEnergy(torch.autograd.Function):
    """Differentiable score(f): forward launches K4 once and keeps d score/d f; backward is a scale."""

    @staticmethod
    def forward(ctx, f, targets, g, l, gs, ls, normalize_f):
        score, _per, _k, grad = energy_fwd_bwd(f.detach(), targets, g, l, gs, ls, normalize_f)
        ctx.save_for_backward(grad)
        return score

    @staticmethod
    def backward(ctx, g_score):
        (grad,) = ctx.saved_tensors
        return grad * g_score, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# K1 / K2 / K3 / K3'  prototype construction
# ------------------------------------------------------------------------------------------------
def sort_by_class(labels: torch.Tensor, num_classes: int):
    """Index plumbing (torch): stable sort of the labels -> (perm [N] i64, class_off [C+1] i64)."""
    labels = _req(labels, "labels").to(torch.int64)
    if labels.numel() and (int(labels.min()) < 0 or int(labels.max()) >= num_classes):
        raise DistDiffError(f"label out of range [0, {num_classes})")
    _, perm = torch.sort(labels, stable=True)
    counts = torch.bincount(labels, minlength=num_classes)
    off = torch.zeros(num_classes + 1, dtype=torch.int64, device=labels.device)
    off[1:] = torch.cumsum(counts, 0)
    return perm.contiguous(), off


def proto_workspace(D: int, C_: int, K: int, device) -> torch.Tensor:
    nbytes = _lib.lib().dd_proto_workspace_bytes(D, C_, K)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def rownorm_classsum(feat: torch.Tensor, perm: Optional[torch.Tensor], class_off: torch.Tensor,
                     ws: Optional[torch.Tensor] = None):
    """K1: -> (feat_sorted [N,D] f32 normalised, class_sum [C,D] f64, class_cnt [C] i64).  dataloader.py:677-707."""
    feat = _req(feat, "features", torch.float32)
    N, D = feat.shape
    class_off = _req(class_off, "class_off", torch.int64)
    C_ = class_off.numel() - 1
    if perm is not None:
        perm = _req(perm, "perm", torch.int64)
    ws = proto_workspace(D, C_, 1, feat.device) if ws is None else ws
    out = torch.empty_like(feat)
    csum = torch.empty(C_, D, dtype=torch.float64, device=feat.device)
    ccnt = torch.empty(C_, dtype=torch.int64, device=feat.device)
    _call("dd_rownorm_classsum", 2 * N * D * 4 + N * 8, 2,
          _ptr(feat), _ptr(perm), _ptr(class_off), N, D, C_, _ptr(out), _ptr(csum),
                                         _ptr(ccnt), _ptr(ws), ws.numel(), _stream())
    return out, csum, ccnt


def class_mean(sum_: torch.Tensor, cnt: torch.Tensor, want_unit: bool = True):
    """K2: fp32 means (+ unit-norm copy) of [.., D] f64 sums / i64 counts.  dataloader.py:707; generate_data.py:1115-1116."""
    sum_ = _req(sum_, "sum", torch.float64)
    cnt = _req(cnt, "cnt", torch.int64)
    D = sum_.shape[-1]
    R = sum_.numel() // D
    mean = torch.empty(sum_.shape, dtype=torch.float32, device=sum_.device)
    unit = torch.empty_like(mean) if want_unit else None
    _call("dd_class_mean", sum_.numel() * 8 + mean.numel() * 4 * (2 if want_unit else 1), 1,
          _ptr(sum_), _ptr(cnt), R, D, _ptr(mean), _ptr(unit), _stream())
    return mean, unit


def normalize_rows(t: torch.Tensor) -> torch.Tensor:
    """rows / ||rows||  (generate_data.py:1115-1116, 1121-1122)."""
    t = _req(t, "prototypes", torch.float32)
    D = t.shape[-1]
    out = torch.empty_like(t)
    _call("dd_normalize_rows", 2 * t.numel() * 4, 1,
          _ptr(t), t.numel() // D, D, _ptr(out), _stream())
    return out


def kmeans_seed(x_sorted: torch.Tensor, row_idx: torch.Tensor, out=None):
    """Seed rows -> (sum [.., D] f64 = the row or 0, cnt [..] i64 = 1 or 0); ``out=(sum, cnt)`` writes in place."""
    x_sorted = _req(x_sorted, "x_sorted", torch.float32)
    row_idx = _req(row_idx, "row_idx", torch.int64)
    D = x_sorted.shape[1]
    R = row_idx.numel()
    if out is None:
        s = torch.empty(*row_idx.shape, D, dtype=torch.float64, device=x_sorted.device)
        c = torch.empty(row_idx.shape, dtype=torch.int64, device=x_sorted.device)
    else:
        s, c = _req(out[0], "sum", torch.float64), _req(out[1], "cnt", torch.int64)
        if s.numel() != R * D or c.numel() != R or s.data_ptr() != out[0].data_ptr():
            raise DistDiffError("kmeans_seed: out buffers must be contiguous [R, D] f64 / [R] i64")
    _call("dd_kmeans_seed", R * D * 12, 1,
          _ptr(x_sorted), _ptr(row_idx), R, D, _ptr(s), _ptr(c), _stream())
    return s, c


def kmeans_update(sum_: torch.Tensor, cnt: torch.Tensor, centroid: torch.Tensor, cnorm: torch.Tensor) -> None:
    Cn, K, D = centroid.shape
    _call("dd_kmeans_update", centroid.numel() * 12, 1,
          _ptr(_req(sum_, "sum", torch.float64)), _ptr(_req(cnt, "cnt", torch.int64)), Cn, K, D,
                                      _ptr(centroid), _ptr(cnorm), _stream())


class KMeansBuffers:
    """Caller-owned buffers of one k-means problem (allocated once, reused every Lloyd iteration).  With ``arena`` (a
    PeerArena) the buffers every rank exchanges live in the IPC-shared arena, at the same offsets on all ranks."""

    def __init__(self, N: int, D: int, C_: int, K: int, device, arena=None):
        self.arena = arena
        if arena is None:
            self.centroid = torch.zeros(C_, K, D, dtype=torch.float32, device=device)
            self.cnorm = torch.zeros(C_, K, dtype=torch.float32, device=device)
            self.sum = torch.empty(C_, K, D, dtype=torch.float64, device=device)
            self.cnt = torch.empty(C_, K, dtype=torch.int64, device=device)
            self.gcnt = None
        else:   # arena memory is zero-initialised by dd_peer_create
            self.sum = arena.carve((C_, K, D), torch.float64)
            self.cnt = arena.carve((C_, K), torch.int64)
            self.centroid = arena.carve((C_, K, D), torch.float32)
            self.cnorm = arena.carve((C_, K), torch.float32)
            self.gcnt = arena.carve((C_, K), torch.int64)
        self.assign = torch.empty(N, dtype=torch.int32, device=device)
        self.inertia = torch.empty(1, dtype=torch.float64, device=device)
        self.ws = proto_workspace(D, C_, K, device)

    @staticmethod
    def arena_bytes(D: int, C_: int, K: int, world: int = 1) -> int:
        """Arena bytes behind the header: the five exchanged buffers + the exchange kernel's inbox at the tail."""
        return int(_lib.lib().dd_peer_arena_bytes(C_ * K, D, world))


def kmeans_assign_accum(x_sorted: torch.Tensor, class_off: torch.Tensor, buf: KMeansBuffers, want_inertia: bool = False,
                        mma: bool = False) -> None:
    """K3: one assignment + accumulation pass; results land in buf.assign / buf.sum / buf.cnt (and buf.inertia when
    ``want_inertia``: that variant reduces one more value per row, so fewer rows share a reduction round).
    ``mma``: K = 4..10 with the dot products on the tensor cores (split-fp16 mma.sync; same assignments, measured slower
    than the default FMA-pipe kernel -- kept as the measured alternative)."""
    x_sorted = _req(x_sorted, "x_sorted", torch.float32)
    N, D = x_sorted.shape
    Cn, K, _ = buf.centroid.shape
    _call("dd_kmeans_assign_accum", N * D * 4 + 2 * N * 4, 2 if want_inertia else 1,
          _ptr(x_sorted), _ptr(class_off), N, D, Cn, K, _ptr(buf.centroid), _ptr(buf.cnorm),
                                            _ptr(buf.assign), _ptr(buf.sum), _ptr(buf.cnt), _ptr(buf.inertia) if want_inertia else None, _ptr(buf.ws),
                                            buf.ws.numel(), int(bool(mma)), _stream())


def kmeans_lloyd(x_sorted: torch.Tensor, class_off: torch.Tensor, buf: KMeansBuffers, iters: int, comm=None) -> None:
    """``iters`` Lloyd iterations launched back to back from C (dd_kmeans_lloyd): K3 pass + exchange per iteration.
    Exchange: the fused peer kernel when the buffers live in a PeerArena, else NCCL all-reduce (``comm``: an ops.Comm)
    + update, else the update alone (single GPU)."""
    x_sorted = _req(x_sorted, "x_sorted", torch.float32)
    N, D = x_sorted.shape
    Cn, K, _ = buf.centroid.shape
    peer = buf.arena.ctx if buf.arena is not None else None
    nccl = comm.handle if (comm is not None and peer is None) else None
    _call("dd_kmeans_lloyd", int(iters) * (N * D * 4 + 2 * N * 4), 2 * int(iters),
          _ptr(x_sorted), _ptr(class_off), N, D, Cn, K, _ptr(buf.centroid), _ptr(buf.cnorm), _ptr(buf.assign), _ptr(buf.sum),
          _ptr(buf.cnt), _ptr(buf.gcnt), _ptr(buf.ws), buf.ws.numel(), nccl, peer, int(iters), _stream())


def agglo_average(x_sorted: torch.Tensor, class_off: torch.Tensor, K: int, max_class_size: int):
    """K3': -> (labels [N] i32 in class-sorted order, sum [C,K,D] f64, cnt [C,K] i64, status [C] i32)."""
    x_sorted = _req(x_sorted, "x_sorted", torch.float32)
    class_off = _req(class_off, "class_off", torch.int64)
    N, D = x_sorted.shape
    Cn = class_off.numel() - 1
    nbytes = _lib.lib().dd_agglo_workspace_bytes(max(int(max_class_size), 1), Cn)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x_sorted.device)
    labels = torch.empty(N, dtype=torch.int32, device=x_sorted.device)
    s = torch.empty(Cn, K, D, dtype=torch.float64, device=x_sorted.device)
    c = torch.empty(Cn, K, dtype=torch.int64, device=x_sorted.device)
    status = torch.empty(Cn, dtype=torch.int32, device=x_sorted.device)
    _call("dd_agglo_average", N * D * 4, 1,
          _ptr(x_sorted), _ptr(class_off), Cn, D, K, max(int(max_class_size), 1), _ptr(labels),
                                      _ptr(s), _ptr(c), _ptr(status), _ptr(ws), nbytes, _stream())
    return labels, s, c, status


# ------------------------------------------------------------------------------------------------
# NCCL communicator through the C ABI
# ------------------------------------------------------------------------------------------------
class Comm:
    """One NCCL communicator per process; the 128-byte unique id travels through ``exchange`` (a callable
    rank0_bytes -> bytes on every rank, e.g. a torch.distributed broadcast_object_list)."""

    def __init__(self, rank: int, world: int, exchange):
        self.rank, self.world = rank, world
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            check(_lib.lib().dd_comm_unique_id(buf), "dd_comm_unique_id")
        uid = exchange(bytes(buf))
        buf2 = (C.c_ubyte * 128).from_buffer_copy(uid)
        handle = C.c_void_p()
        check(_lib.lib().dd_comm_init(rank, world, buf2, C.byref(handle)), "dd_comm_init")
        self.handle = handle

    def allreduce(self, sum_: Optional[torch.Tensor], cnt: Optional[torch.Tensor]) -> None:
        ns = 0 if sum_ is None else sum_.numel()
        nc = 0 if cnt is None else cnt.numel()
        if sum_ is not None:
            _req(sum_, "sum", torch.float64)
        if cnt is not None:
            _req(cnt, "cnt", torch.int64)
        check(_lib.lib().dd_comm_allreduce(self.handle, _ptr(sum_), ns, _ptr(cnt), nc, _stream()), "dd_comm_allreduce")

    def close(self) -> None:
        if self.handle:
            check(_lib.lib().dd_comm_destroy(self.handle), "dd_comm_destroy")
            self.handle = None


# ------------------------------------------------------------------------------------------------
# fused centroid exchange over NVLink peer memory (csrc/dd_peer.cu)
# ------------------------------------------------------------------------------------------------
class _DeviceSpan:
    """Raw device memory as a `__cuda_array_interface__` object so torch can view it (the arena is cudaMalloc'ed by
    the library because it has to be IPC-exportable; torch's caching allocator sub-allocates and cannot be)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class PeerArena:
    """One IPC-shared arena per rank; ``allgather_bytes(b) -> [bytes per rank]`` is the host exchange of the handles."""

    def __init__(self, rank: int, world: int, nbytes: int, allgather_bytes, device):
        self.rank, self.world, self.device = rank, world, torch.device(device)
        L = _lib.lib()
        self.header = int(L.dd_peer_header_bytes())
        self.nbytes = int(nbytes) + self.header
        handle = (C.c_ubyte * 64)()
        ctx = C.c_void_p()
        check(L.dd_peer_create(rank, world, self.nbytes, C.byref(ctx), handle), "dd_peer_create")
        self.ctx = ctx
        blobs = allgather_bytes(bytes(handle))
        if len(blobs) != world or any(len(b) != 64 for b in blobs):
            raise DistDiffError("PeerArena: the handle exchange must return one 64-byte handle per rank")
        check(L.dd_peer_connect(self.ctx, b"".join(blobs)), "dd_peer_connect")
        self._span = _DeviceSpan(int(L.dd_peer_local(self.ctx)), self.nbytes)
        self.bytes_view = torch.as_tensor(self._span, device=self.device)
        self._cursor = self.header

    def reset(self) -> None:
        """Forget the carved buffers (the next problem re-carves from the start; flags and epoch carry on)."""
        self._cursor = self.header

    @property
    def capacity(self) -> int:
        return self.nbytes - self.header

    def carve(self, shape, dtype) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        nb = n * torch.empty((), dtype=dtype).element_size()
        off = (self._cursor + 255) // 256 * 256
        if off + nb > self.nbytes:
            raise DistDiffError("PeerArena: out of space")
        self._cursor = off + nb
        t = self.bytes_view[off:off + nb].view(dtype).view(*shape)
        t._dd_arena_offset = off
        return t

    def kmeans_exchange(self, sum_, cnt, centroid, cnorm, gcnt) -> None:
        R, D = cnt.numel(), sum_.shape[-1]
        _call("dd_peer_kmeans_exchange", 0, 1, self.ctx, sum_._dd_arena_offset, cnt._dd_arena_offset, centroid._dd_arena_offset,
              cnorm._dd_arena_offset, gcnt._dd_arena_offset, R, D, _stream())

    def status(self) -> int:
        v = C.c_int(0)
        check(_lib.lib().dd_peer_status(self.ctx, _stream(), C.byref(v)), "dd_peer_status")
        return int(v.value)

    def timing(self):
        """Phase boundaries (us since kernel start) of the last exchange on this rank: see dd_peer_timing."""
        v = (C.c_double * 5)()
        check(_lib.lib().dd_peer_timing(self.ctx, _stream(), v), "dd_peer_timing")
        return [float(x) for x in v]

    def close(self) -> None:
        if self.ctx:
            self.bytes_view = None
            check(_lib.lib().dd_peer_destroy(self.ctx), "dd_peer_destroy")
            self.ctx = None
