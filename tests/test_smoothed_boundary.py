"""Shape (pde_opt/numerics/shapes.py), the smoothed-boundary equations (cahn_hilliard.py:203-289, allen_cahn.py:87-159)
and the explicit solvers that integrate them."""
import numpy as np
import pytest
import torch

from oracle import pde_oracle as O
from pde_opt_b200 import Domain
from pde_opt_b200.shapes import Shape
from pde_opt_b200.solvers import Dopri5, Euler, ODETerm


def _disc(n, r):
    x = np.arange(n) - n / 2 + 0.5
    X, Y = np.meshgrid(x, x, indexing="ij")
    return (X**2 + Y**2 < r**2).astype(np.float64)


def test_shape_smoothing_and_clipping():
    s = Shape(_disc(48, 14), dx=(1.0, 1.0), smooth_epsilon=3.0, smooth_curvature=0.008, smooth_dt=0.01, smooth_tf=5.0)
    assert s.smooth.shape == (48, 48) and s.smooth.min() >= 0.001 and s.smooth.max() <= 1.0  # shapes.py:37-38
    assert s.smooth[24, 24] == 1.0 and s.smooth[0, 0] == 0.001
    inter = (s.smooth > 0.05) & (s.smooth < 0.95)
    assert inter.sum() > 0  # a diffuse interface replaced the jump
    # the smoothing flow is the reference's expression: its RHS vanishes on a constant field
    assert np.allclose(s.flow_rhs(np.zeros((8, 8))), 0.0) and np.allclose(s.flow_rhs(np.ones((8, 8))), 0.0)


def test_graph_laplacian_of_mask():
    m = np.zeros((4, 5)); m[1:3, 1:4] = 1  # a 2 x 3 block of pixels: 6 nodes, 7 edges
    s = Shape.__new__(Shape); s.binary = m
    L, ids = s.laplacian_from_mask()
    assert L.shape == (6, 6) and (ids >= 0).sum() == 6
    assert np.allclose(L.sum(1), 0) and L.diagonal().sum() == 14 and (L != L.T).nnz == 0
    Lp, _ = s.laplacian_from_mask(periodic=True)
    assert Lp.diagonal().sum() == 14  # the block does not touch the boundary: same graph
    s.get_shape_modes(N=3)
    assert s.shape_basis.shape == (4, 5, 3) and abs(s.shape_basis_evals[0]) < 1e-6  # constant mode first


def test_dopri5_order_and_error_estimate():
    """dy/dt = -y + sin t on CPU tensors: fifth-order convergence, error estimate of the size of the true local error."""
    term = ODETerm(lambda t, y, args: -y + np.sin(t))

    def run(solver, n):
        y, ts = torch.tensor([1.0], dtype=torch.float64), np.linspace(0.0, 1.0, n + 1)
        for a, b in zip(ts[:-1], ts[1:]):
            y = solver.step(term, np.float64(a), np.float64(b), y)[0]
        return float(y)

    exact = 1.5 * np.exp(-1.0) + 0.5 * (np.sin(1.0) - np.cos(1.0))
    e1, e2 = abs(run(Dopri5(), 4) - exact), abs(run(Dopri5(), 8) - exact)
    assert 4.2 < np.log2(e1 / e2) < 5.8  # float32 time stamps inside step() limit the observed order slightly
    ee1, ee2 = abs(run(Euler(), 64) - exact), abs(run(Euler(), 128) - exact)
    assert 0.9 < np.log2(ee1 / ee2) < 1.1
    _, err, *_ = Dopri5().step(term, 0.0, 0.25, torch.tensor([1.0], dtype=torch.float64))
    assert err is not None and 1e-9 < abs(float(err)) < 1e-4


def _sbm_problem(n):
    H = 0.01
    shape = Shape(_disc(n, 0.3 * n), dx=(1.0, 1.0), smooth_epsilon=3.0, smooth_curvature=0.008, smooth_dt=0.01, smooth_tf=4.0)
    dom = Domain((n, n), ((-n * H / 2, n * H / 2),) * 2, "dimensionless", shape)
    f_t = lambda c: c * torch.log(c) + (1 - c) * torch.log(1 - c) + 3.0 * c * (1 - c) + 0.059
    mu_t = lambda c: torch.log(c / (1 - c)) + 3.0 * (1 - 2 * c)
    mob_t = lambda c: (1 - c) * c
    f_n = lambda c: c * np.log(c) + (1 - c) * np.log(1 - c) + c.dtype.type(3.0) * c * (1 - c) + c.dtype.type(0.059)
    mu_n = lambda c: np.log(c / (1 - c)) + c.dtype.type(3.0) * (1 - 2 * c)
    mob_n = lambda c: (1 - c) * c
    return dom, shape, (f_t, mu_t, mob_t), (f_n, mu_n, mob_n)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ch", "ac"])
def test_smoothed_boundary_rhs_matches_oracle(kind):
    from pde_opt_b200.equations import AllenCahn2DSmoothedBoundary, CahnHilliard2DSmoothedBoundary

    n, KAPPA = 128, 0.002
    dom, shape, (f_t, mu_t, mob_t), (f_n, mu_n, mob_n) = _sbm_problem(n)
    theta, flux = (lambda t: np.pi / 3 + 0.1 * t), (lambda t: 0.2 + t)
    B = 3
    y0 = np.stack([np.clip(0.5 + 0.1 * np.random.default_rng(i).normal(size=(n, n)), 0.05, 0.95) for i in range(B)]).astype(np.float32)
    psi = shape.smooth.astype(np.float32)
    if kind == "ch":
        eq = CahnHilliard2DSmoothedBoundary(dom, KAPPA, f_t, mu_t, mob_t, theta, flux)
        side = np.zeros((n, n), np.float32); side[:50, :] = 1
        want = [O.sbm_rhs_ch(y0[b], 0.5, psi, dom.dx, KAPPA, f_n, mu_n, mob_n, theta, flux, side) for b in range(B)]
    else:
        eq = AllenCahn2DSmoothedBoundary(dom, KAPPA, f_t, mu_t, mob_t, theta)
        side = np.zeros((n, n), np.float32); side[:, :100] = 1
        want = [O.sbm_rhs_ac(y0[b], 0.5, psi, dom.dx, KAPPA, f_n, mu_n, mob_n, theta, side) for b in range(B)]
    got = eq.rhs(torch.from_numpy(y0).cuda(), 0.5).cpu().numpy()
    for b in range(B):
        # float64 evaluation of the same expression is the yardstick for float32 rounding (the flux differences cancel strongly)
        ref64 = (O.sbm_rhs_ch(y0[b].astype(np.float64), 0.5, psi.astype(np.float64), dom.dx, KAPPA, f_n, mu_n, mob_n, theta, flux, side.astype(np.float64))
                 if kind == "ch" else
                 O.sbm_rhs_ac(y0[b].astype(np.float64), 0.5, psi.astype(np.float64), dom.dx, KAPPA, f_n, mu_n, mob_n, theta, side.astype(np.float64)))
        e_gpu = np.linalg.norm(got[b] - ref64) / np.linalg.norm(ref64)
        e_orc = np.linalg.norm(want[b] - ref64) / np.linalg.norm(ref64)
        assert e_gpu <= max(2.0 * e_orc, 1e-5), (kind, b, e_gpu, e_orc)
    assert torch.allclose(eq.rhs(torch.from_numpy(y0[1]).cuda(), 0.5), torch.from_numpy(got[1]).cuda())  # [nx, ny] input


@pytest.mark.gpu
def test_smoothed_boundary_solve_with_dopri5_and_pid():
    """The notebook's call pattern (solving_pde_smoothed_boundary.ipynb): PDEModel(AllenCahn2DSmoothedBoundary, domain,
    <explicit solver>).solve(..., stepsize_controller=PIDController(rtol, atol)); the adaptive solution matches a
    fine constant-step solution, and the field stays inside (0, 1)."""
    from pde_opt_b200.equations import AllenCahn2DSmoothedBoundary
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.stepsize import PIDController

    n, KAPPA = 128, 0.002
    dom, shape, (f_t, mu_t, mob_t), _ = _sbm_problem(n)
    model = PDEModel(AllenCahn2DSmoothedBoundary, dom, Dopri5)
    params = {"kappa": KAPPA, "f": f_t, "mu": mu_t, "R": mob_t, "theta": lambda t: np.pi / 2.0, "derivs": "fd"}
    y0 = torch.from_numpy(np.clip(0.5 + 0.1 * np.random.default_rng(0).normal(size=(n, n)), 0.05, 0.95).astype(np.float32)).cuda()
    ts = np.array([0.0, 0.01, 0.02], np.float32)
    ada = model.solve(params, y0, ts, dt0=1e-4, stepsize_controller=PIDController(rtol=1e-5, atol=1e-7))
    fine = model.solve(params, y0, ts, dt0=1e-4)
    assert model.last_stats["accepted"] > 0
    assert torch.isfinite(ada).all() and float(ada.min()) > 0 and float(ada.max()) < 1
    assert float((ada[-1] - fine[-1]).norm() / fine[-1].norm()) <= 1e-4
    assert float((fine[-1] - y0).norm()) > 1e-3  # something happened
