"""GPU parity of the generic-size fused SIFS kernel (power-of-two grids other than 128x128):
BASELINE config 1 (Allen-Cahn 64x64, single env, 1000 steps) and the reference's own 256x1
Cahn-Hilliard known-answer test run through PDEModel.solve on the GPU."""
import numpy as np
import pytest

from oracle import pde_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_config1_allen_cahn_64_single_env_1000_steps():
    """BASELINE.json configs[0]: Allen-Cahn 2D 64x64 periodic, semi-implicit spectral, single
    env, 1000 steps (dx=0.01, kappa=0.002, mu=c^3-c, R=1, dt=5e-6, A=1)."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AllenCahn2DPeriodic
    from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral

    n, h, kappa, dt = 64, 0.01, 0.002, 5e-6
    dom = Domain((n, n), ((-n * h / 2, n * h / 2),) * 2, "dimensionless")
    eq = AllenCahn2DPeriodic(dom, kappa, lambda c: c**3 - c, lambda c: 0 * c + 1.0)
    solver = SemiImplicitFourierSpectral(1.0, eq.fourier_symbol, eq.fft, eq.ifft)
    u0 = (0.01 * np.random.default_rng(0).normal(size=(n, n))).astype(np.float32)
    oeq = O.AllenCahn2DPeriodic(O.Domain((n, n), ((-n * h / 2, n * h / 2),) * 2), kappa, O.mu_double_well,
                                lambda c: np.ones_like(c), "fd", np.float32)
    times = O.constant_step_schedule(0.0, 1000 * dt, dt, np.float32)
    assert abs(len(times) - 1 - 1000) <= 1
    # one step
    y1 = solver.rollout(ODETerm(eq), times[:2], torch.from_numpy(u0).cuda()).cpu().numpy()
    ref1 = O.sifs_step(oeq.rhs, u0, times[0], times[1], 1.0, oeq.fourier_symbol)
    assert rel_l2(y1, ref1) <= 1e-5
    assert rel_l2(y1 - u0, ref1 - u0) <= 1e-3
    # 1000 steps
    yN = solver.rollout(ODETerm(eq), times, torch.from_numpy(u0).cuda()).cpu().numpy()
    y = u0
    for a, b in zip(times[:-1], times[1:]):
        y = O.sifs_step(oeq.rhs, y, a, b, 1.0, oeq.fourier_symbol)
    assert rel_l2(yN, y) <= 1e-3


@pytest.mark.parametrize("shape", [(64, 64), (32, 32), (64, 128), (32, 16), (256, 1)])
def test_generic_cahn_hilliard_shapes(shape):
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    nx, ny = shape
    h, kappa = 0.01, 0.002
    odom = O.Domain((nx, ny), ((0.0, nx * h), (0.0, ny * h)))
    oeq = O.CahnHilliardPeriodic(odom, kappa, lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, "fd", np.float32)
    plan = SifsPlan("ch2d", nx, ny, (0.0, 0.0), (h, h), kappa, ("log", (3.0,)), ("degenerate", ()))
    sym = torch.from_numpy(fold_symbol(oeq.fourier_symbol, 0.5)).cuda()
    y0 = np.stack([np.clip(0.5 + 0.01 * np.random.default_rng(s).normal(size=(nx, ny)), 0, 1) for s in range(3)]).astype(np.float32)
    dts = [1e-6 * (1 + 1e-3 * k) for k in range(6)]
    rew = torch.empty((3, 2), device="cuda")
    got = plan.step(torch.from_numpy(y0).cuda(), dts, sym, reward=rew).cpu().numpy()
    for b in range(3):
        y, t = y0[b], np.float32(0)
        for d in dts:
            y = O.sifs_step(oeq.rhs, y, t, t + np.float32(d), 0.5, oeq.fourier_symbol)
            t = t + np.float32(d)
        assert rel_l2(got[b], y) <= 1e-5
        assert rel_l2(got[b] - y0[b], y - y0[b]) <= 2e-3
        np.testing.assert_allclose(rew[b, 1].item(), got[b].astype(np.float64).var(), rtol=1e-3)
    f = plan.rhs(torch.from_numpy(y0).cuda()).cpu().numpy()
    for b in range(3):
        assert rel_l2(f[b], oeq.rhs(y0[b])) <= 5e-4


def test_reference_kat_ch_tanh_profile_on_gpu():
    """reference tests/test_solvers.py:208-248 (test_1d_cahn_hilliard_pde_model) through our
    PDEModel.solve on the GPU, float32: 256x1, kappa=.002, mu=c^3-c, D=1, A=.5, dt=5e-5, t=10,
    SaveAt(linspace(0,10,200)), vs tanh(x/sqrt(2 kappa)) on the middle half, rtol=atol=1e-3."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard2DPeriodic
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    Nx, Ny = 256, 1
    Lx, Ly = 0.01 * Nx, 0.01 * Ny
    dom = Domain((Nx, Ny), ((-Lx / 2, Lx / 2), (-Ly / 2, Ly / 2)), "dimensionless")
    kappa = 0.002
    model = PDEModel(CahnHilliard2DPeriodic, dom, SemiImplicitFourierSpectral)
    u0 = np.ones((Nx, Ny), np.float32)
    u0[: Nx // 2] = -1.0
    ys = model.solve({"kappa": kappa, "mu": lambda c: c**3 - c, "D": lambda c: np.ones_like(c), "derivs": "fd"},
                     torch.from_numpy(u0).cuda(), np.linspace(0.0, 10.0, 200), {"A": 0.5}, dt0=0.00005, max_steps=1000000)
    analytic = np.tanh(dom.axes()[0] / np.sqrt(2 * kappa))
    np.testing.assert_allclose(ys[-1].cpu().numpy().squeeze()[Nx // 4 : 3 * Nx // 4], analytic[Nx // 4 : 3 * Nx // 4],
                               rtol=1e-3, atol=1e-3)
