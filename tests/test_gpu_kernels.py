"""GPU parity tests: every kernel of libdistdiff_sm100.so, called through the C ABI (ctypes, via
distdiff_b200.ops), against the CPU oracle on the same seeded inputs and against the committed golden
vectors produced by the reference's own function bodies.

Tolerances (north_star): fp32 elementwise kernels are BIT-EXACT against the oracle's fp32 op sequence;
reductions (energy, prototypes) within 1e-5 relative (prototypes) / 1e-5 (energy score & gradient, fp32
accumulation order differs); indices (k*, assignments, agglomerative labels) exact; fp16/bf16 storage:
one rounding of the fp32 result (<= 1 ulp of the storage type, stated per test).
"""
import numpy as np
import pytest
import torch

from oracle import ddim, energy, guidance, prototypes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from distdiff_b200 import ops as _ops
    return _ops


def _g(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------- K5
@pytest.mark.parametrize("shape", [(1, 4, 64, 64), (2, 4, 64, 64), (3, 4, 8, 8), (2, 3, 5, 7)])
@pytest.mark.parametrize("cfg", [True, False])
@pytest.mark.parametrize("with_grad", [False, True])
def test_cfg_ddim_fwd_fp32_bitexact(ops, cuda_device, shape, cfg, with_grad):
    g = _g(1)
    x = torch.randn(shape, generator=g)
    npred = torch.randn((2 * shape[0] if cfg else shape[0],) + shape[1:], generator=g)
    grad = torch.randn(shape, generator=g) if with_grad else None
    for t in (981, 381, 1):
        a_t, a_prev = ddim.alpha_pair(t)
        if cfg:
            rp, r0 = ddim.cfg_ddim_step(npred, x, 7.5, a_t, a_prev, grad, 10.0)
        else:
            rp, r0 = ddim.ddim_step(npred, x, a_t, a_prev)
            if grad is not None:
                rp = rp - 10.0 * grad
        p, x0 = ops.cfg_ddim_step(npred.to(cuda_device), x.to(cuda_device), 7.5, float(a_t), float(a_prev), cfg=cfg,
                                  grad=None if grad is None else grad.to(cuda_device), rho=10.0)
        assert torch.equal(p.cpu(), rp), (t, (p.cpu() - rp).abs().max())
        assert torch.equal(x0.cpu(), r0)


def test_cfg_ddim_fwd_golden(ops, cuda_device, golden):
    for name in ("guidance_small", "guidance_d2048", "guidance_b1"):
        c = golden[name]
        a_t, a_prev = ddim.alpha_pair(c["denoise"]["t"])
        p, x0 = ops.cfg_ddim_step(c["denoise"]["noise_pred"].to(cuda_device), c["latents"].to(cuda_device),
                                  c["args"]["guidance_scale"], float(a_t), float(a_prev))
        assert torch.equal(p.cpu(), c["denoise"]["prev"]) and torch.equal(x0.cpu(), c["denoise"]["x0"])


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2 ** -10), (torch.bfloat16, 2 ** -7)])
def test_cfg_ddim_fwd_half(ops, cuda_device, dtype, tol):
    """16-bit storage: fp32 arithmetic, one rounding -> within 1 ulp (relative 2^-10 fp16 / 2^-7 bf16) of the
    fp32 oracle evaluated on the same (already rounded) inputs."""
    g = _g(2)
    x = torch.randn(4, 4, 64, 64, generator=g).to(dtype)
    npred = torch.randn(8, 4, 64, 64, generator=g).to(dtype)
    a_t, a_prev = ddim.alpha_pair(481)
    rp, r0 = ddim.cfg_ddim_step(npred.float(), x.float(), 7.5, a_t, a_prev)
    p, x0 = ops.cfg_ddim_step(npred.to(cuda_device), x.to(cuda_device), 7.5, float(a_t), float(a_prev))
    assert p.dtype == dtype
    assert torch.allclose(p.float().cpu(), rp, rtol=tol, atol=tol * 1e-2)
    assert torch.allclose(x0.float().cpu(), r0, rtol=tol, atol=tol * 1e-2)
    # one rounding of an fp32 result that is itself within 1 fp32 ulp of the oracle: storage values are equal except
    # where that ulp straddles a rounding boundary (rare)
    assert (p.cpu() != rp.to(dtype)).float().mean() < 1e-3 and (x0.cpu() != r0.to(dtype)).float().mean() < 1e-3


@pytest.mark.parametrize("cfg", [True, False])
def test_cfg_ddim_bwd_vs_autograd(ops, cuda_device, cfg):
    g = _g(3)
    B = 2
    x = torch.randn(B, 4, 16, 16, generator=g, dtype=torch.float64, requires_grad=True)
    npred = torch.randn(2 * B if cfg else B, 4, 16, 16, generator=g, dtype=torch.float64, requires_grad=True)
    wp = torch.randn(B, 4, 16, 16, generator=g, dtype=torch.float64)
    w0 = torch.randn(B, 4, 16, 16, generator=g, dtype=torch.float64)
    a_t, a_prev = (v.double() for v in ddim.alpha_pair(381))
    if cfg:
        rp, r0 = ddim.cfg_ddim_step(npred, x, 7.5, a_t, a_prev)
    else:
        rp, r0 = ddim.ddim_step(npred, x, a_t, a_prev)
    ((rp * wp).sum() + (r0 * w0).sum()).backward()

    xc = x.detach().float().to(cuda_device).requires_grad_(True)
    nc = npred.detach().float().to(cuda_device).requires_grad_(True)
    p, x0 = ops.CfgDdimStep.apply(nc, xc, 7.5, float(a_t), float(a_prev), cfg)
    ((p * wp.float().to(cuda_device)).sum() + (x0 * w0.float().to(cuda_device)).sum()).backward()
    assert torch.allclose(xc.grad.cpu().double(), x.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(nc.grad.cpu().double(), npred.grad, rtol=1e-5, atol=1e-4)
    # only x0 used (direct_guidance): g_prev is None
    xc.grad = None; nc.grad = None
    p, x0 = ops.CfgDdimStep.apply(nc, xc, 7.5, float(a_t), float(a_prev), cfg)
    (x0 * w0.float().to(cuda_device)).sum().backward()
    x.grad = None; npred.grad = None
    if cfg:
        _, r0 = ddim.cfg_ddim_step(npred, x, 7.5, a_t, a_prev)
    else:
        _, r0 = ddim.ddim_step(npred, x, a_t, a_prev)
    (r0 * w0).sum().backward()
    assert torch.allclose(xc.grad.cpu().double(), x.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(nc.grad.cpu().double(), npred.grad, rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------- K6 / K7
@pytest.mark.parametrize("shape", [(2, 4, 64, 64), (1, 4, 8, 8), (3, 5, 3, 3)])
@pytest.mark.parametrize("radius", [-1.0, 0.2, 0.8])
def test_affine_project_fp32_bitexact(ops, cuda_device, shape, radius):
    g = _g(4)
    x = torch.randn(shape, generator=g)
    a = torch.rand(shape[0], shape[1], 1, 1, generator=g)
    b = torch.randn(shape[0], shape[1], 1, 1, generator=g)
    ref = guidance.affine_project(x, a, b, radius if radius >= 0 else None)
    if radius >= 0:  # literally the reference sequence: transform then in-place masked clamp
        lit = x * (1 + a) + b
        guidance.linfball_proj(x.clone(), radius, lit, in_place=True)
        assert torch.equal(lit, ref)
    y = ops.affine_project(x.to(cuda_device), a.to(cuda_device), b.to(cuda_device), radius)
    assert torch.equal(y.cpu(), ref)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("with_center", [False, True])
def test_affine_project_large_batch_16bit(ops, cuda_device, dtype, with_center):
    """Large 16-bit batches take the two-rows-per-CTA variant (odd row count: the last CTA has one row); it must be
    bit-identical to the one-row variant that small batches use, and one rounding away from the fp32 result."""
    g = torch.Generator(device=cuda_device).manual_seed(6)
    B, Cc = 481, 5                                              # 2405 rows >= 16*148: two-row path with an odd tail
    x = torch.randn(B, Cc, 64, 64, generator=g, device=cuda_device).to(dtype)
    c = (x.float() + 0.3 * torch.randn(B, Cc, 64, 64, generator=g, device=cuda_device)).to(dtype) if with_center else None
    a = torch.rand(B, Cc, 1, 1, generator=g, device=cuda_device)
    b = torch.randn(B, Cc, 1, 1, generator=g, device=cuda_device)
    big = ops.affine_project(x, a, b, 0.2, center=c)
    small = torch.cat([ops.affine_project(x[i:i + 2], a[i:i + 2], b[i:i + 2], 0.2, center=None if c is None else c[i:i + 2])
                       for i in range(0, B, 2)])
    assert torch.equal(big, small)
    ref32 = ops.affine_project(x.float(), a, b, 0.2, center=None if c is None else c.float())
    assert torch.equal(big, ref32.to(dtype))


def test_linfball_golden(ops, cuda_device, golden):
    c = golden["linfball"]
    B, Cc = c["t"].shape[:2]
    z = torch.zeros(B, Cc, 1, 1)
    y = ops.affine_project(c["t"].to(cuda_device), z.to(cuda_device), z.to(cuda_device), c["radius"],
                           center=c["center"].to(cuda_device))
    assert torch.equal(y.cpu(), c["out"])


def test_affine_bwd_vs_autograd(ops, cuda_device):
    g = _g(5)
    x = torch.randn(3, 4, 64, 64, generator=g, dtype=torch.float64)
    a = torch.rand(3, 4, 1, 1, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn(3, 4, 1, 1, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(3, 4, 64, 64, generator=g, dtype=torch.float64)
    ((x * (1 + a) + b) * w).sum().backward()
    xc = x.float().to(cuda_device).requires_grad_(True)
    ac = a.detach().float().to(cuda_device).requires_grad_(True)
    bc = b.detach().float().to(cuda_device).requires_grad_(True)
    (ops.ChannelAffine.apply(xc, ac, bc) * w.float().to(cuda_device)).sum().backward()
    assert torch.allclose(ac.grad.cpu().double(), a.grad, rtol=1e-5, atol=1e-3)
    assert torch.allclose(bc.grad.cpu().double(), b.grad, rtol=1e-5, atol=1e-3)
    assert torch.allclose(xc.grad.cpu().double(), w * (1 + a.detach()), rtol=1e-6, atol=1e-6)
    # determinism: two runs are bit-identical
    ac2 = a.detach().float().to(cuda_device).requires_grad_(True)
    bc2 = b.detach().float().to(cuda_device).requires_grad_(True)
    (ops.ChannelAffine.apply(xc.detach(), ac2, bc2) * w.float().to(cuda_device)).sum().backward()
    assert torch.equal(ac2.grad, ac.grad) and torch.equal(bc2.grad, bc.grad)


@pytest.mark.parametrize("shape", [(2, 4, 64, 64), (1, 3, 5, 7)])
def test_add_noise_bitexact(ops, cuda_device, shape):
    g = _g(6)
    x = torch.randn(shape, generator=g); n = torch.randn(shape, generator=g)
    for t in (981, 481, 1):
        a_t = ddim.alphas_cumprod()[t]
        out = ops.add_noise(x.to(cuda_device), n.to(cuda_device), float(a_t))
        assert torch.equal(out.cpu(), ddim.add_noise(x, n, a_t))


# ------------------------------------------------------------------------------------------- K4
def _energy_case(B, C, K, D, seed, dup=False):
    g = _g(seed)
    f = torch.randn(B, D, generator=g)
    gp = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=-1)
    lp = torch.nn.functional.normalize(torch.randn(C, K, D, generator=g), dim=-1)
    if dup and K > 1:
        lp[:, 1] = lp[:, 0]  # exact argmax ties -> first max must win
    y = torch.randint(0, C, (B,), generator=g).tolist()
    return f, gp, lp, y


@pytest.mark.parametrize("B,C,K,D", [(1, 5, 3, 64), (2, 100, 3, 2048), (16, 100, 3, 2048), (7, 10, 10, 1280),
                                     (5, 4, 1, 512), (1024, 100, 3, 2048), (2000, 50, 10, 2048), (1500, 7, 5, 512),
                                     (64, 1000, 3, 2048)])   # last: ImageNet-scale tables, BASELINE configs[4]
@pytest.mark.parametrize("normalize_f", [False, True])
@pytest.mark.parametrize("mode", ["sample", "tile_cta", "tile_pair"])   # one CTA per sample / class-bucketed kernels (large-B path): thread groups, warp pairs
def test_energy_vs_oracle(ops, cuda_device, B, C, K, D, normalize_f, mode):
    f, gp, lp, y = _energy_case(B, C, K, D, 10 + B)
    for G, L in ((gp, lp), (gp, None), (None, lp)):
        s_ref, per_ref, k_ref, g_ref = energy.energy_fwd_bwd(f.numpy(), y, None if G is None else G.numpy(),
                                                             None if L is None else L.numpy(), 0.7, 1.3, normalize_f)
        m = "tile" if mode == "tile_pair" and (G is None or L is None) else mode   # the warp-pair kernel needs both tables
        score, per, kstar, grad = ops.energy_fwd_bwd(f.to(cuda_device), y, None if G is None else G.to(cuda_device),
                                                     None if L is None else L.to(cuda_device), 0.7, 1.3, normalize_f, mode=m)
        assert abs(float(score) - float(s_ref)) <= 1e-5 * abs(float(s_ref))
        assert np.allclose(per.cpu().numpy(), per_ref, rtol=1e-5, atol=1e-6)
        if L is not None:
            # k* exact, except where the oracle's own top-2 dots are closer than fp32 resolution (documented tie)
            fn = f.double() / f.double().norm(dim=-1, keepdim=True) if normalize_f else f.double()
            dots = torch.einsum("bd,bkd->bk", fn, L.double()[y])
            top2 = dots.topk(min(2, K), dim=-1).values
            clear = (top2[:, 0] - top2[:, -1]).abs() > 1e-5 * dots.abs().max() if K > 1 else torch.ones(B, dtype=torch.bool)
            assert np.array_equal(kstar.cpu().numpy()[clear.numpy()], k_ref[clear.numpy()])
            assert clear.float().mean() > 0.99
        gn = np.linalg.norm(g_ref)
        assert np.linalg.norm(grad.cpu().numpy() - g_ref) <= 1e-5 * gn


def test_energy_vs_autograd_of_literal_restatement(ops, cuda_device):
    f, gp, lp, y = _energy_case(6, 9, 3, 2048, 77)
    for normalize_f in (False, True):
        fr = f.clone().requires_grad_(True)
        s = energy.energy_score(fr, y, gp, lp, 1.0, 1.0, normalize_f=normalize_f)
        (gr,) = torch.autograd.grad(s, fr)
        fc = f.to(cuda_device).requires_grad_(True)
        sc = ops.PrototypeEnergy.apply(fc, y, gp.to(cuda_device), lp.to(cuda_device), 1.0, 1.0, normalize_f)
        (2.0 * sc).backward()
        assert abs(float(sc) - float(s)) < 1e-5 * abs(float(s))
        assert torch.allclose(fc.grad.cpu(), 2.0 * gr, rtol=1e-4, atol=1e-7)


def test_energy_ties_and_zero_distance(ops, cuda_device):
    f, gp, lp, y = _energy_case(8, 6, 4, 256, 5, dup=True)
    _, _, kstar, _ = ops.energy_fwd_bwd(f.to(cuda_device), y, gp.to(cuda_device), lp.to(cuda_device), 1.0, 1.0, False)
    _, _, k_ref, _ = energy.energy_fwd_bwd(f.numpy(), y, gp.numpy(), lp.numpy(), 1.0, 1.0, False)
    assert np.array_equal(kstar.cpu().numpy(), k_ref)
    assert not np.any(kstar.cpu().numpy() == 1)  # duplicate of index 0 never wins
    # f exactly on its class prototype: distance 0 -> gradient 0 (torch.norm sub-gradient), no NaN
    y2 = [1, 2, 3]
    f2 = gp[y2].clone()
    score, per, _, grad = ops.energy_fwd_bwd(f2.to(cuda_device), y2, gp.to(cuda_device), None, 1.0, 1.0, False)
    assert float(score) == 0.0 and torch.count_nonzero(grad) == 0 and torch.isfinite(grad).all()


@pytest.mark.parametrize("K", [1, 3, 5, 8, 10, 16])
def test_energy_tile_matches_sample_kernel(ops, cuda_device, K):
    """auto mode at B >= 16 x SMs takes the class-tiled kernel; it must agree with the per-sample kernel on the same
    inputs (ragged classes, an empty class, unsorted targets), for every register-tile instantiation."""
    B, C, D = 5000, 37, 2048
    f, gp, lp, y = _energy_case(B, C, K, D, 100 + K)
    y = [v if v != 5 else 6 for v in y]          # class 5 stays empty
    args = (f.to(cuda_device), y, gp.to(cuda_device), lp.to(cuda_device), 1.0, 0.5)
    for nf in (False, True):
        a = ops.energy_fwd_bwd(*args, nf, mode="sample")
        b = ops.energy_fwd_bwd(*args, nf, mode="auto")
        c = ops.energy_fwd_bwd(*args, nf, mode="tile")
        assert torch.equal(b[0], c[0]) and torch.equal(b[3], c[3])        # auto == tile, bit for bit
        for b in (c, ops.energy_fwd_bwd(*args, nf, mode="tile_cta"), *([ops.energy_fwd_bwd(*args, nf, mode="tile_pair")] if K <= 10 else [])):
            assert abs(float(a[0]) - float(b[0])) <= 2e-6 * abs(float(a[0]))
            assert torch.allclose(a[1], b[1], rtol=2e-6, atol=1e-7)
            assert (a[2] != b[2]).float().mean() < 1e-3                     # only fp32-level argmax ties may differ
            assert (a[3] - b[3]).norm() <= 2e-6 * a[3].norm()


@pytest.mark.parametrize("C", [1000, 5000])
def test_energy_class_bucketing_many_classes(ops, cuda_device, C):
    """The class bucketing of the large-batch kernels: shared-memory histogram (C + 2 <= 4096) and the warp-aggregated global
    atomics beyond that; both large-batch kernels must agree with the per-sample kernel when most class runs hold 1-5 samples."""
    B, K, D = 20000, 3, 512
    f, gp, lp, y = _energy_case(B, C, K, D, 900 + C)
    args = (f.to(cuda_device), y, gp.to(cuda_device), lp.to(cuda_device), 1.0, 1.0, True)
    a = ops.energy_fwd_bwd(*args, mode="sample")
    for mode in ("tile_cta", "tile_pair"):
        b = ops.energy_fwd_bwd(*args, mode=mode)
        assert abs(float(a[0]) - float(b[0])) <= 2e-6 * abs(float(a[0]))
        assert torch.allclose(a[1], b[1], rtol=2e-6, atol=1e-7)
        assert (a[2] != b[2]).float().mean() < 1e-3
        assert (a[3] - b[3]).norm() <= 2e-6 * a[3].norm()


@pytest.mark.parametrize("mode", ["tile_pair", "tile_cta"])
@pytest.mark.parametrize("normalize_f", [False, True])
def test_energy_tile_small_distances_take_the_exact_pass(ops, cuda_device, normalize_f, mode):
    """Samples on / next to their prototypes: the tile kernel's distance-from-dots expansion would lose accuracy there,
    so the batch falls back to direct sums -- distances, score and gradient still match the oracle, exact zeros included."""
    B, C, K, D = 257, 9, 4, 2048
    f, gp, lp, y = _energy_case(B, C, K, D, 21)
    f = torch.nn.functional.normalize(f, dim=-1) if not normalize_f else f
    yt = torch.tensor(y)
    scale = 3.7 if normalize_f else 1.0                       # direct mode normalises f first
    f[0] = gp[yt[0]] * scale                                   # exactly on the class prototype
    f[1] = lp[yt[1], 2] * scale                                # exactly on a group prototype
    f[2:40] = (gp[yt[2:40]] + 1e-3 * torch.randn(38, D, generator=_g(5))) * scale
    f[40:80] = (lp[yt[40:80], 1] + 3e-2 * torch.randn(40, D, generator=_g(6)) / D ** 0.5) * scale
    s_ref, per_ref, k_ref, g_ref = energy.energy_fwd_bwd(f.numpy(), y, gp.numpy(), lp.numpy(), 1.0, 1.0, normalize_f)
    score, per, kstar, grad = ops.energy_fwd_bwd(f.to(cuda_device), y, gp.to(cuda_device), lp.to(cuda_device), 1.0, 1.0,
                                                 normalize_f, mode=mode)
    assert abs(float(score) - float(s_ref)) <= 1e-5 * abs(float(s_ref))
    assert np.allclose(per.cpu().numpy()[2:], per_ref[2:], rtol=2e-5, atol=1e-6)
    assert np.allclose(per.cpu().numpy()[:2], per_ref[:2], rtol=1e-5, atol=2e-6)     # ~0 entries: fp32 rounding of fn
    assert np.array_equal(kstar.cpu().numpy()[80:], k_ref[80:])
    assert torch.isfinite(grad).all()
    # the gradient of a ~zero distance is a unit vector of rounding noise in BOTH implementations: compare the rest
    gn = np.linalg.norm(g_ref[80:])
    assert np.linalg.norm(grad.cpu().numpy()[80:] - g_ref[80:]) <= 1e-5 * gn
    if not normalize_f:
        f2 = gp[yt[:8]].clone()
        sc, pr, _, gr = ops.energy_fwd_bwd(f2.to(cuda_device), y[:8], gp.to(cuda_device), None, 1.0, 1.0, False, mode="tile")
        assert float(sc) == 0.0 and torch.count_nonzero(gr) == 0


def test_energy_tile_bad_target_is_poisoned(ops, cuda_device):
    B, C, K, D = 300, 6, 3, 512
    f, gp, lp, y = _energy_case(B, C, K, D, 3)
    yt = torch.tensor(y, dtype=torch.int64)
    yt[7] = C
    yt[200] = -1
    for mode in ("sample", "tile_cta", "tile_pair"):
        score, per, kstar, grad = ops.energy_fwd_bwd(f.to(cuda_device), yt.to(cuda_device), gp.to(cuda_device), lp.to(cuda_device),
                                                     1.0, 1.0, False, mode=mode)
        assert torch.isnan(score) and torch.isnan(grad[7]).all() and torch.isnan(grad[200]).all()
        ok = torch.ones(B, dtype=torch.bool); ok[7] = ok[200] = False
        assert torch.isfinite(grad.cpu()[ok]).all() and torch.isfinite(per.cpu()[ok]).all()


def test_energy_deterministic_and_bad_target(ops, cuda_device):
    f, gp, lp, y = _energy_case(3000, 20, 3, 2048, 9)
    a = ops.energy_fwd_bwd(f.to(cuda_device), y, gp.to(cuda_device), lp.to(cuda_device), 1.0, 1.0, True)
    b = ops.energy_fwd_bwd(f.to(cuda_device), y, gp.to(cuda_device), lp.to(cuda_device), 1.0, 1.0, True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3])
    from distdiff_b200._lib import DistDiffError
    with pytest.raises(DistDiffError):
        ops.energy_fwd_bwd(f[:2].to(cuda_device), [0, 20], gp.to(cuda_device), None, 1.0, 1.0)
    with pytest.raises(DistDiffError):
        ops.energy_fwd_bwd(f[:2], [0, 1], gp, None, 1.0, 1.0)  # CPU tensors: no fallback


# ------------------------------------------------------------------------------------------- K1 / K2
def _features(N, C, D, seed, ragged=True):
    rng = np.random.default_rng(seed)
    if ragged:
        labels = rng.integers(0, C, size=N)
        labels[:C] = np.arange(C) if N >= C else labels[:C]
    else:
        labels = np.arange(N) % C
    centers = rng.normal(size=(C, 4, D)).astype(np.float32)
    feats = centers[labels, rng.integers(0, 4, size=N)] * 1.5 + rng.normal(size=(N, D)).astype(np.float32)
    return feats.astype(np.float32), labels.astype(np.int64)


@pytest.mark.parametrize("N,C,D", [(300, 10, 64), (3000, 100, 2048), (50, 7, 1280), (5000, 3, 512), (40, 60, 2048)])
def test_rownorm_classsum_vs_oracle(ops, cuda_device, N, C, D):
    feats, labels = _features(N, C, D, N + C, ragged=True)
    if N < C:
        labels = np.random.default_rng(0).integers(0, C, size=N)  # some classes empty
    ft = torch.from_numpy(feats).to(cuda_device)
    perm, off = ops.sort_by_class(torch.from_numpy(labels).to(cuda_device), C)
    xs, csum, ccnt = ops.rownorm_classsum(ft, perm, off)
    ref = prototypes.l2_normalize_rows(feats)
    order = np.argsort(labels, kind="stable")
    assert np.array_equal(perm.cpu().numpy(), order)
    assert np.allclose(xs.cpu().numpy(), ref[order], rtol=1e-6, atol=1e-9)  # norm summation order: <= a few fp32 ulps
    assert np.array_equal(ccnt.cpu().numpy(), np.bincount(labels, minlength=C))
    ref_sum = np.zeros((C, D)); np.add.at(ref_sum, labels, ref.astype(np.float64))
    assert np.allclose(csum.cpu().numpy(), ref_sum, rtol=1e-6, atol=1e-7)
    mean, unit = ops.class_mean(csum, ccnt)
    nz = np.bincount(labels, minlength=C) > 0
    ref_mean = np.stack([ref[labels == c].mean(0) if nz[c] else np.zeros(D, np.float32) for c in range(C)])
    assert np.allclose(mean.cpu().numpy(), ref_mean, rtol=1e-5, atol=1e-8)
    g_unit, _ = prototypes.normalize_prototypes(ref_mean[nz], ref_mean[nz][:, None, :])
    assert np.allclose(unit.cpu().numpy()[nz], g_unit, rtol=1e-5, atol=1e-8)
    # bit-reproducible
    xs2, csum2, _ = ops.rownorm_classsum(ft, perm, off)
    assert torch.equal(xs, xs2) and torch.equal(csum, csum2)


@pytest.mark.parametrize("name", ["proto_caltech_like", "proto_d2048", "proto_k5"])
def test_class_means_golden(ops, cuda_device, golden, name):
    c = golden[name]
    C = int(c["labels"].max()) + 1
    perm, off = ops.sort_by_class(c["labels"].to(cuda_device), C)
    xs, csum, ccnt = ops.rownorm_classsum(c["features"].to(cuda_device), perm, off)
    mean, _ = ops.class_mean(csum, ccnt, want_unit=False)
    ref = c["global_prototypes"].numpy()
    assert np.linalg.norm(mean.cpu().numpy() - ref, axis=-1).max() <= 1e-5 * np.linalg.norm(ref, axis=-1).min()
    assert np.allclose(mean.cpu().numpy(), ref, rtol=1e-5, atol=1e-8)


# ------------------------------------------------------------------------------------------- K3'
@pytest.mark.parametrize("name", ["proto_caltech_like", "proto_d2048", "proto_k5"])
def test_agglomerative_golden(ops, cuda_device, golden, name):
    c = golden[name]
    K = c["K"]
    labels_np = c["labels"].numpy()
    C = int(labels_np.max()) + 1
    perm, off = ops.sort_by_class(c["labels"].to(cuda_device), C)
    xs, _, ccnt = ops.rownorm_classsum(c["features"].to(cuda_device), perm, off)
    lab, s, n, status = ops.agglo_average(xs, off, K, int(ccnt.max()))
    assert int(status.abs().sum()) == 0
    local, _ = ops.class_mean(s, n, want_unit=False)
    ref = c["local_prototypes"].numpy()
    assert np.allclose(local.cpu().numpy(), ref, rtol=1e-5, atol=1e-8)     # sklearn's cluster numbering too
    # labels == sklearn's, class by class
    fn = prototypes.l2_normalize_rows(c["features"].numpy())
    _, _, ref_labels = prototypes.extract_prototype_from_features(fn, labels_np.tolist(), K)
    offc = off.cpu().numpy(); labc = lab.cpu().numpy()
    for ci in range(C):
        assert np.array_equal(labc[offc[ci]:offc[ci + 1]], ref_labels[ci])


@pytest.mark.parametrize("N,C,D,K", [(600, 12, 64, 3), (900, 6, 256, 4), (400, 5, 2048, 3), (1200, 4, 32, 10), (300, 20, 128, 1)])
def test_agglomerative_vs_sklearn_random(ops, cuda_device, N, C, D, K):
    feats, labels = _features(N, C, D, 3 * N + K, ragged=True)
    fn = prototypes.l2_normalize_rows(feats)
    gl, lc, ref_labels = prototypes.extract_prototype_from_features(fn, labels.tolist(), K)
    perm, off = ops.sort_by_class(torch.from_numpy(labels).to(cuda_device), C)
    xs, _, ccnt = ops.rownorm_classsum(torch.from_numpy(feats).to(cuda_device), perm, off)
    lab, s, n, status = ops.agglo_average(xs, off, K, int(ccnt.max()))
    assert int(status.abs().sum()) == 0
    offc = off.cpu().numpy(); labc = lab.cpu().numpy()
    for ci in range(C):
        assert np.array_equal(labc[offc[ci]:offc[ci + 1]], ref_labels[ci]), ci
    local, _ = ops.class_mean(s, n, want_unit=False)
    assert np.allclose(local.cpu().numpy(), lc, rtol=1e-5, atol=1e-7)


def test_agglomerative_error_status(ops, cuda_device):
    # class 0: 1 sample (sklearn: "at least 2 samples"), class 1: 2 samples < K=3, class 2: fine
    feats = np.random.default_rng(1).normal(size=(8, 64)).astype(np.float32)
    labels = np.array([0, 1, 1, 2, 2, 2, 2, 2])
    perm, off = ops.sort_by_class(torch.from_numpy(labels).to(cuda_device), 3)
    xs, _, _ = ops.rownorm_classsum(torch.from_numpy(feats).to(cuda_device), perm, off)
    _, _, _, status = ops.agglo_average(xs, off, 3, 5)
    assert status.cpu().tolist() == [1, 2, 0]


# ------------------------------------------------------------------------------------------- K3
def _kmeans_problem(ops, dev, N, C, D, K, seed):
    feats, labels = _features(N, C, D, seed, ragged=True)
    perm, off = ops.sort_by_class(torch.from_numpy(labels).to(dev), C)
    xs, _, ccnt = ops.rownorm_classsum(torch.from_numpy(feats).to(dev), perm, off)
    return feats, labels, xs, off, ccnt


@pytest.mark.parametrize("N,C,D,K", [(4000, 10, 2048, 3), (3000, 100, 2048, 3), (6000, 7, 512, 10), (2500, 5, 1280, 5),
                                     (5000, 4, 2048, 8), (700, 3, 64, 1), (3000, 6, 2048, 7), (2000, 3, 256, 12)])
def test_kmeans_one_iteration_vs_oracle(ops, cuda_device, N, C, D, K):
    feats, labels, xs, off, ccnt = _kmeans_problem(ops, cuda_device, N, C, D, K, N + K)
    xs_np = xs.cpu().numpy(); offc = off.cpu().numpy()
    buf = ops.KMeansBuffers(N, D, C, K, cuda_device)
    rng = np.random.default_rng(K)
    mu = np.stack([xs_np[offc[c]:offc[c + 1]][rng.choice(offc[c + 1] - offc[c], K, replace=False)] for c in range(C)])
    mu = (mu + 0.05 * rng.normal(size=mu.shape)).astype(np.float32)
    buf.centroid.copy_(torch.from_numpy(mu))
    buf.cnorm.copy_(torch.from_numpy((mu.astype(np.float64) ** 2).sum(-1).astype(np.float32)))
    ops.kmeans_assign_accum(xs, off, buf, want_inertia=True)
    assign = buf.assign.cpu().numpy()
    inert_ref = 0.0
    clear_all = np.ones(N, bool)
    for c in range(C):
        Xc = xs_np[offc[c]:offc[c + 1]]
        a_ref, s = prototypes.kmeans_assign(Xc, mu[c])
        srt = np.sort(s, axis=-1)
        clear = (srt[:, 1] - srt[:, 0] > 1e-5) if K > 1 else np.ones(len(Xc), bool)   # documented tie: top-2 gap <= 1e-5
        a = assign[offc[c]:offc[c + 1]]
        clear_all[offc[c]:offc[c + 1]] = clear
        assert np.array_equal(a[clear], a_ref[clear]), c
        assert clear.mean() > 0.995
        sums_ref, cnt_ref = prototypes.kmeans_sums(Xc, a, K)        # sums for the GPU's own assignment
        assert np.array_equal(buf.cnt[c].cpu().numpy(), cnt_ref)
        scale = np.abs(sums_ref).max() + 1e-30
        assert np.abs(buf.sum[c].cpu().numpy() - sums_ref).max() <= 2e-6 * scale
        inert_ref += (((Xc.astype(np.float64) - mu[c][a].astype(np.float64)) ** 2).sum())
    assert abs(float(buf.inertia) - inert_ref) <= 1e-4 * inert_ref + 1e-6
    # reproducible bit for bit
    s1 = buf.sum.clone(); a1 = buf.assign.clone(); c1 = buf.cnt.clone()
    ops.kmeans_assign_accum(xs, off, buf, want_inertia=True)
    assert torch.equal(s1, buf.sum) and torch.equal(a1, buf.assign)
    # the inertia-free call runs, for K = 4..10, the cluster-paired FMA kernel or (mma=True, D % 256 == 0) the tensor-core
    # kernel (split-fp16 MMA) -- other summation orders: same assignment except on the documented ties, counts and sums
    # exact / to fp32 rounding for ITS OWN assignment, reproducible
    for mma in (False, True):
        ops.kmeans_assign_accum(xs, off, buf, want_inertia=False, mma=mma)
        a2 = buf.assign.cpu().numpy()
        assert np.array_equal(a2[clear_all], assign[clear_all]), mma
        for c in range(C):
            Xc = xs_np[offc[c]:offc[c + 1]]
            sums_ref, cnt_ref = prototypes.kmeans_sums(Xc, a2[offc[c]:offc[c + 1]], K)
            assert np.array_equal(buf.cnt[c].cpu().numpy(), cnt_ref)
            assert np.abs(buf.sum[c].cpu().numpy() - sums_ref).max() <= 2e-6 * (np.abs(sums_ref).max() + 1e-30)
        s2 = buf.sum.clone(); a2t = buf.assign.clone()
        ops.kmeans_assign_accum(xs, off, buf, want_inertia=False, mma=mma)
        assert torch.equal(s2, buf.sum) and torch.equal(a2t, buf.assign)
    s1 = buf.sum.clone()
    # update kernel
    old = buf.centroid.clone()
    ops.kmeans_update(buf.sum, buf.cnt, buf.centroid, buf.cnorm)
    cnt = buf.cnt.cpu().numpy()
    new_ref = np.where(cnt[..., None] > 0, (buf.sum.cpu().numpy() / np.maximum(cnt, 1)[..., None]).astype(np.float32),
                       old.cpu().numpy())
    assert np.array_equal(buf.centroid.cpu().numpy(), new_ref)
    assert np.allclose(buf.cnorm.cpu().numpy(), (new_ref.astype(np.float64) ** 2).sum(-1), rtol=1e-6)


def test_kmeans_tiny_and_empty(ops, cuda_device):
    # fewer rows than CTAs, an empty class in the middle, N == 0
    feats = np.random.default_rng(2).normal(size=(9, 64)).astype(np.float32)
    labels = np.array([0, 0, 0, 2, 2, 2, 2, 3, 3])
    perm, off = ops.sort_by_class(torch.from_numpy(labels).to(cuda_device), 4)
    xs, csum, ccnt = ops.rownorm_classsum(torch.from_numpy(feats).to(cuda_device), perm, off)
    assert ccnt.cpu().tolist() == [3, 0, 4, 2] and float(csum[1].abs().sum()) == 0.0
    buf = ops.KMeansBuffers(9, 64, 4, 2, cuda_device)
    buf.centroid.copy_(torch.randn(4, 2, 64)); buf.cnorm.copy_((buf.centroid.double() ** 2).sum(-1).float())
    ops.kmeans_assign_accum(xs, off, buf)
    assert buf.cnt.sum(-1).cpu().tolist() == [3, 0, 4, 2]
    ops.kmeans_assign_accum(xs, off, buf, want_inertia=True)
    assert buf.cnt.sum(-1).cpu().tolist() == [3, 0, 4, 2]
    xs0 = torch.empty(0, 64, device=cuda_device)
    off0 = torch.zeros(5, dtype=torch.int64, device=cuda_device)
    buf0 = ops.KMeansBuffers(0, 64, 4, 2, cuda_device)
    ops.kmeans_assign_accum(xs0, off0, buf0, want_inertia=True)
    assert int(buf0.cnt.sum()) == 0 and float(buf0.inertia) == 0.0


# ------------------------------------------------------------------------------------------- K8 bicubic resize
_RS_CASES = [((2, 3, 512, 512), (224, 224)), ((1, 3, 100, 77), (37, 50)), ((1, 2, 64, 64), (64, 64)),
             ((1, 1, 48, 96), (56, 120)), ((3, 1, 130, 40), (30, 9)), ((1, 3, 256, 256), (224, 224)),
             ((2, 3, 64, 64), (224, 224)), ((1, 1, 5, 7), (3, 2))]


@pytest.mark.parametrize("shape,size", _RS_CASES)
def test_bicubic_resize_fwd_fp32(ops, cuda_device, shape, size):
    from oracle import resize
    x = torch.randn(shape, generator=_g(11))
    got = ops.bicubic_resize(x.to(cuda_device), size).cpu()
    ref = resize.interpolate_bicubic(x, size)                  # the reference's call (generate_data.py:704), CPU fp32
    ref64 = resize.bicubic_numpy(x.numpy(), size)              # ATen's arithmetic restated, fp64 accumulation
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 1e-5 * scale      # fp32: accumulation order / FMA contraction only
    assert np.abs(got.numpy() - ref64).max() <= 1e-5 * scale
    # same op on the GPU through ATen: the arithmetic order is ATen's, so this is (near) bit-exact
    aten = torch.nn.functional.interpolate(x.to(cuda_device), size=size, mode="bicubic").cpu()
    assert float((got - aten).abs().max()) <= 2e-6 * scale


@pytest.mark.parametrize("dtype,ulp", [(torch.float16, 2.0 ** -10), (torch.bfloat16, 2.0 ** -7)])
def test_bicubic_resize_fwd_half(ops, cuda_device, dtype, ulp):
    from oracle import resize
    x = torch.randn((2, 3, 512, 512), generator=_g(12)).to(dtype)
    got = ops.bicubic_resize(x.to(cuda_device), (224, 224))
    assert got.dtype == dtype
    ref = resize.bicubic_numpy(x.float().numpy(), (224, 224))  # exact result of the stored inputs
    err = np.abs(got.float().cpu().numpy() - ref)
    assert (err <= ulp * np.maximum(np.abs(ref), 1e-3) + 1e-6).all()   # one rounding to the storage type


@pytest.mark.parametrize("shape,size", _RS_CASES)
def test_bicubic_resize_bwd(ops, cuda_device, shape, size):
    from oracle import resize
    g = torch.randn((shape[0], shape[1]) + size, generator=_g(13))
    got = ops.bicubic_resize_bwd(g.to(cuda_device), shape[2:])
    ref = resize.bicubic_backward_numpy(g.numpy(), shape[2:])
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    again = ops.bicubic_resize_bwd(g.to(cuda_device), shape[2:])
    assert torch.equal(got, again)                              # gather, fixed order: bit-reproducible


def test_bicubic_tma_staging_matches_plain_staging(ops, cuda_device, monkeypatch):
    """fp32 forward / backward and 16-bit forward: the tensor-map (TMA) staged kernels and the load/store staged ones they
    replace give the same bits (same two passes, same order), edge tiles included; DD_K8_NO_TMA is the library's development switch."""
    x = torch.randn(3, 3, 512, 512, generator=_g(31)).to(cuda_device)
    g = torch.randn(3, 3, 224, 224, generator=_g(32)).to(cuda_device)
    small = torch.randn(2, 1, 64, 48, generator=_g(33)).to(cuda_device)
    outs = []
    for no_tma in ("", "1"):
        if no_tma:
            monkeypatch.setenv("DD_K8_NO_TMA", no_tma)
        outs.append((ops.bicubic_resize(x, (224, 224)), ops.bicubic_resize_bwd(g, (512, 512)), ops.bicubic_resize(small, (20, 30)),
                     ops.bicubic_resize(x.half(), (224, 224)), ops.bicubic_resize(x.bfloat16(), (224, 224)),
                     ops.bicubic_resize(small[..., :40].contiguous().half(), (30, 25))))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_bicubic_resize_autograd(ops, cuda_device):
    x = torch.randn((2, 3, 96, 80), generator=_g(14))
    xd = x.to(cuda_device).requires_grad_(True)
    w = torch.randn((2, 3, 42, 35), generator=_g(15))
    (ops.BicubicResize.apply(xd, (42, 35)) * w.to(cuda_device)).sum().backward()
    xr = x.clone().requires_grad_(True)   # fp32 autograd of the reference's call (fp64 would move the tap positions)
    (torch.nn.functional.interpolate(xr, size=(42, 35), mode="bicubic") * w).sum().backward()
    assert float((xd.grad.cpu() - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())


def test_bicubic_resize_errors(ops, cuda_device):
    from distdiff_b200._lib import DistDiffError
    with pytest.raises(DistDiffError):
        ops.bicubic_resize(torch.randn(1, 3, 8, 8), (4, 4))                     # CPU tensor: no fallback
    with pytest.raises(DistDiffError):
        ops.bicubic_resize(torch.randn(1, 1, 512, 512, device=cuda_device), (16, 16))      # 32x down-scaling: region > smem


# ------------------------------------------------------------------------------------------- K9 image -> uint8
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (1, 1, 6, 6), (1, 4, 10, 14), (1, 3, 512, 512)])
@pytest.mark.parametrize("denorm", [True, False])
def test_image_to_uint8_bitexact(ops, cuda_device, dtype, shape, denorm):
    from oracle import image as o_img
    x = torch.randn(shape, generator=_g(21)) * (0.9 if denorm else 0.4) + (0.0 if denorm else 0.5)
    flat = x.view(-1)
    # exact half-way and boundary values of the quantiser
    special = torch.tensor([-1.0, 1.0, 0.0, -1.5, 1.5, 1 / 255, 0.5 / 255, 127.5 / 255, 254.5 / 255, 2 * 0.5 / 255 - 1, 2 * 100.5 / 255 - 1])
    flat[:special.numel()] = special
    x = x.to(dtype)
    got = ops.image_to_uint8(x.to(cuda_device), denormalize=denorm).cpu().numpy()
    ref = o_img.decode_to_uint8(x, do_denormalize=denorm)       # the reference's eager ops in the image dtype, on the CPU
    assert got.shape == ref.shape and got.dtype == np.uint8
    assert np.array_equal(got, ref)


def test_png_writer_matches_save_image(ops, cuda_device, tmp_path):
    from torchvision.utils import save_image
    from distdiff_b200.expand import AsyncPngWriter
    from oracle import image as o_img
    x = (torch.randn((3, 3, 64, 48), generator=_g(22)) * 0.8).half()
    ref_paths = [str(tmp_path / f"ref_{i}.png") for i in range(3)]
    img = o_img.denormalize(x)                                   # generate_data.py:1227
    for i, p in enumerate(ref_paths):
        save_image([img[i]], p)                                  # generate_data.py:1234
    ours = [str(tmp_path / "cls" / f"ours_{i}.png") for i in range(3)]
    with AsyncPngWriter(workers=2, depth=1) as w:
        w.submit(ops.image_to_uint8(x.to(cuda_device)), ours)
    for a, b in zip(ref_paths, ours):
        assert open(a, "rb").read() == open(b, "rb").read()      # byte-identical files
    assert not [f for f in (tmp_path / "cls").iterdir() if ".tmp." in f.name]
