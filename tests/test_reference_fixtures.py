"""Golden vectors of the TRUE reference (jax + diffrax), written by tools/make_reference_fixtures.py on a box where
the reference is importable.  This image has no jax, so the files are absent here and every test skips — parity at the
north-star tolerances stays "unpinned" until they are committed.  When present they pin the
NumPy oracle (CPU tests) and the CUDA path (-m gpu) at: relative L2 <= 1e-5 after 1 and 16 steps, <= 1e-3 after 1000."""
import glob
import os

import numpy as np
import pytest

from oracle import pde_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
H, KAPPA = 0.01, 0.002


def _load(name):
    path = os.path.join(GOLD, f"ref_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{os.path.basename(path)} not generated (needs the jax reference: tools/make_reference_fixtures.py)")
    return np.load(path)


def _steps(fx):
    return sorted(int(k[2:]) for k in fx.files if k.startswith("y_"))


def _tol(k):
    return 1e-5 if k <= 16 else 1e-3  # north star: 1e-5 after one step, 1e-3 after 1000


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))


def _dom(*n):
    return O.Domain(tuple(n), tuple((-k * H / 2, k * H / 2) for k in n))


def _oracle_sifs(eq, y0, dt, k, A):
    ts = np.arange(k + 1, dtype=np.float32) * np.float32(dt)
    y = y0
    for i in range(k):
        y = O.sifs_step(eq.rhs, y, ts[i], ts[i + 1], A, eq.fourier_symbol)
    return y


def test_fixture_inventory():
    """At least say which fixtures exist (the suite must not silently pass on an empty set without saying so)."""
    found = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "ref_*.npz")))
    if not found:
        pytest.skip("no reference fixtures committed: parity against the jax reference is unpinned")


@pytest.mark.parametrize("name", ["c1_ac64", "c2_ch128", "c2_ch128_ps", "c5_ch3d32"])
def test_oracle_matches_reference_sifs(name):
    fx = _load(name)
    y0, dt, A = fx["y0"], float(fx["dt"]), float(fx["A"])
    if name == "c1_ac64":
        eq = O.AllenCahn2DPeriodic(_dom(64, 64), KAPPA, O.mu_double_well, lambda c: np.ones_like(c), "fd", np.float32)
    elif name == "c5_ch3d32":
        eq = O.CahnHilliardPeriodic(_dom(32, 32, 32), KAPPA, lambda c: O.mu_log(c, 3.0), lambda c: 0.15 * np.ones_like(c), "fd", np.float32)
    else:
        eq = O.CahnHilliardPeriodic(_dom(128, 128), KAPPA, lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, "fd", np.float32)
    for k in _steps(fx):
        assert _rel(_oracle_sifs(eq, y0, dt, k, A), fx[f"y_{k}"]) <= _tol(k), (name, k)


@pytest.mark.parametrize("name", ["c3_gpe128_imag", "c3_gpe128_real"])
def test_oracle_matches_reference_strang(name):
    fx = _load(name)
    y0, dt, ts_ = fx["y0"], float(fx["dt"]), complex(fx["time_scale"])
    dom = O.Domain((128, 128), ((-7.5, 7.5), (-7.5, 7.5)))
    eq = O.GPE2DTSControl(dom, 3371.7, 0.0, lambda t, x, y: np.zeros_like(x), 1.0, np.float32)
    for k in _steps(fx):
        t = np.arange(k + 1, dtype=np.float32) * np.float32(dt)
        y = y0
        for i in range(k):
            y = O.strang_step(eq.B_terms, y, t[i], t[i + 1], eq.A_term, eq.dx, ts_)
        assert _rel(y, fx[f"y_{k}"]) <= _tol(k), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_ac64", "c2_ch128", "c2_ch128_ps", "c5_ch3d32"])
def test_cuda_matches_reference_sifs(name):
    import torch

    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AllenCahn2DPeriodic, CahnHilliard2DPeriodic, CahnHilliard3DPeriodic
    from pde_opt_b200.functions import ConstantMobility, DegenerateMobility, DoubleWell, LogRegular
    from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral

    fx = _load(name)
    y0, dt, A = fx["y0"], float(fx["dt"]), float(fx["A"])
    n = y0.shape
    dom = Domain(tuple(n), tuple((-k * H / 2, k * H / 2) for k in n), "dimensionless")
    if name == "c1_ac64":
        eq = AllenCahn2DPeriodic(dom, KAPPA, DoubleWell(), ConstantMobility(1.0))
    elif name == "c5_ch3d32":
        eq = CahnHilliard3DPeriodic(dom, KAPPA, LogRegular(3.0), ConstantMobility(0.15))
    else:
        eq = CahnHilliard2DPeriodic(dom, KAPPA, LogRegular(3.0), DegenerateMobility())
    solver = SemiImplicitFourierSpectral(A, eq.fourier_symbol, eq.fft, eq.ifft)
    for k in _steps(fx):
        times = np.arange(k + 1, dtype=np.float32) * np.float32(dt)
        y = solver.rollout(ODETerm(eq), times, torch.from_numpy(y0[None]).cuda())
        assert _rel(y[0].cpu().numpy(), fx[f"y_{k}"]) <= _tol(k), (name, k)


@pytest.mark.gpu
def test_cuda_gradients_match_jax_grad():
    import torch

    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard2DPeriodic
    from pde_opt_b200.functions import ChemicalPotentialLegendrePolynomials, DiffusionLegendrePolynomials
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    fx = _load("grad_ch64")
    dom = Domain((64, 64), ((-0.32, 0.32),) * 2, "dimensionless")
    model = PDEModel(CahnHilliard2DPeriodic, dom, SemiImplicitFourierSpectral)
    mu_c = torch.tensor(fx["mu_coef"], device="cuda", requires_grad=True)
    d_c = torch.tensor(fx["d_coef"], device="cuda", requires_grad=True)
    params = {"kappa": KAPPA, "mu": ChemicalPotentialLegendrePolynomials(mu_c, "log"), "D": DiffusionLegendrePolynomials(d_c)}
    loss = model.mse(params, (torch.from_numpy(fx["y0s"]).cuda(), torch.from_numpy(fx["target"]).cuda()), {"A": float(fx["A"])},
                     fx["ts"], {}, 0.0, dt0=float(fx["dt"]))
    loss.backward()
    assert abs(float(loss) - float(fx["loss"])) <= 1e-4 * abs(float(fx["loss"]))
    for got, want in ((mu_c.grad.cpu().numpy(), fx["g_mu"]), (d_c.grad.cpu().numpy(), fx["g_d"])):
        assert np.linalg.norm(got - want) <= 1e-4 * np.linalg.norm(want)  # north star: gradients to relative 1e-4
