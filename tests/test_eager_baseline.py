"""The benchmark comparator (distdiff_b200/eager_baseline.py: the literal PyTorch-eager sequences the kernels replace) must
compute what the oracle computes -- otherwise "fused vs eager" in bench.py would compare different things.  CPU tensors."""
import numpy as np
import torch

from distdiff_b200 import eager_baseline as eager
from oracle import ddim, energy


def _g(seed):
    return torch.Generator().manual_seed(seed)


def test_eager_cfg_ddim_step_and_add_noise_match_oracle():
    g = _g(0)
    x = torch.randn(2, 4, 8, 8, generator=g)
    npred = torch.randn(4, 4, 8, 8, generator=g)
    grad = torch.randn(2, 4, 8, 8, generator=g)
    a_t, a_prev = ddim.alpha_pair(381)
    rp, r0 = ddim.cfg_ddim_step(npred, x, 7.5, a_t, a_prev, grad=grad, rho=10.0)
    p, x0 = eager.cfg_ddim_step(npred, x, 7.5, float(a_t), float(a_prev), grad=grad, rho=10.0)
    assert torch.allclose(p, rp, rtol=1e-6, atol=1e-6) and torch.allclose(x0, r0, rtol=1e-6, atol=1e-6)
    n = torch.randn(2, 4, 8, 8, generator=g)
    assert torch.allclose(eager.add_noise(x, n, float(a_t)), ddim.add_noise(x, n, a_t), rtol=1e-6, atol=1e-7)


def test_eager_energy_matches_oracle_closed_form():
    g = _g(1)
    f = torch.randn(5, 64, generator=g)
    gp = torch.nn.functional.normalize(torch.randn(7, 64, generator=g), dim=-1)
    lp = torch.nn.functional.normalize(torch.randn(7, 3, 64, generator=g), dim=-1)
    y = [0, 3, 6, 2, 2]
    for nf in (False, True):
        s_ref, _, _, g_ref = energy.energy_fwd_bwd(f.numpy(), y, gp.numpy(), lp.numpy(), 0.7, 1.3, nf)
        s, _, _, gr = eager.energy_fwd_bwd(f, y, gp, lp, 0.7, 1.3, nf)
        assert abs(float(s) - float(s_ref)) < 1e-5 * abs(float(s_ref))
        assert np.linalg.norm(gr.numpy() - g_ref) < 1e-5 * np.linalg.norm(g_ref)


def test_eager_projection_is_the_reference_clamp():
    g = _g(2)
    x = torch.randn(2, 3, 4, 4, generator=g)
    a, b = torch.rand(2, 3, 1, 1, generator=g), torch.randn(2, 3, 1, 1, generator=g)
    y = eager.affine_project(x, a, b, 0.2)
    ref = torch.minimum(torch.maximum(x * (1 + a) + b, x - 0.2), x + 0.2)
    assert torch.equal(y, ref)


def test_swapped_restores_the_product_ops():
    from distdiff_b200 import expand, guidance, ops, scheduler
    with eager.swapped():
        assert guidance.ops is eager and expand.ops is eager and scheduler.ops is eager
    assert guidance.ops is ops and expand.ops is ops and scheduler.ops is ops
