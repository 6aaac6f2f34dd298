"""Pin the CPU oracle against the reference's own known-answer tests and fixtures.

Each test names the reference test it re-runs (same parameters, same tolerance)."""

import os

import numpy as np
import pytest

from oracle import pde_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_ch_tanh_profile_kat():
    """reference tests/test_solvers.py:21-61 (test_1d_cahn_hilliard): 256x1, kappa=.002,
    mu=c^3-c, D=1, SIFS(A=.5), dt=5e-5, t=10, vs tanh(x/sqrt(2 kappa)), rtol=atol=1e-3."""
    Nx, Ny = 256, 1
    Lx, Ly = 0.01 * Nx, 0.01 * Ny
    dom = O.Domain((Nx, Ny), ((-Lx / 2, Lx / 2), (-Ly / 2, Ly / 2)))
    kappa = 0.002
    eq = O.CahnHilliardPeriodic(dom, kappa, O.mu_double_well, lambda c: np.ones_like(c), "fd", np.float64)
    u0 = np.ones((Nx, Ny))
    u0[: Nx // 2] = -1.0
    ys = O.integrate(
        lambda y, a, b: O.sifs_step(eq.rhs, y, a, b, 0.5, eq.fourier_symbol),
        u0, 0.0, 10.0, 0.00005, save_ts=np.linspace(0.0, 10.0, 200),
    )
    analytic = np.tanh(dom.axes()[0] / np.sqrt(2 * kappa))
    np.testing.assert_allclose(
        ys[-1].squeeze()[Nx // 4 : 3 * Nx // 4], analytic[Nx // 4 : 3 * Nx // 4], rtol=1e-3, atol=1e-3
    )


def test_gpe_thomas_fermi_kat():
    """reference tests/test_solvers.py:107-205 (test_2d_gross_pitaevskii): imaginary-time
    Strang splitting to the Thomas-Fermi density, rtol=atol=1e-3."""
    atoms = 5e5
    hbar = 1.05e-34
    omega = 2 * np.pi * 10
    omega_z = np.sqrt(8) * omega
    mass = 3.8175406e-26
    a0 = 5.29177210903e-11
    a_s = 100 * a0
    N = 128
    x_s = np.sqrt(hbar / (mass * omega))
    t_s = 1 / omega
    L_ = 150e-6 / x_s
    k = 4 * np.pi * a_s * atoms * np.sqrt((mass * omega_z) / (2 * np.pi * hbar))
    t1_, dt_ = 0.1 / t_s, 1e-5 / t_s
    dom = O.Domain((N, N), ((-L_ / 2, L_ / 2), (-L_ / 2, L_ / 2)))
    psi0 = O.initialize_Psi(N, width=100) * x_s
    psi0 = psi0 / np.sqrt(np.sum(np.abs(psi0) ** 2) * dom.dx[0] ** 2)
    eq = O.GPE2DTSControl(dom, k, 0.0, lambda t, x, y: 0.0 * x, 1.0, np.float64)
    y0 = np.stack([psi0.real, psi0.imag], -1)
    ys = O.integrate(
        lambda y, a, b: O.strang_step(eq.B_terms, y, a, b, eq.A_term, eq.dx, -1j),
        y0, 0.0, t1_, dt_, save_ts=np.linspace(0.0, t1_, 100),
    )
    X, Y = dom.mesh()
    g = k
    mu = np.sqrt((1.0 * g * np.sqrt(0.5) * np.sqrt(0.5)) / (2.0 * np.pi))
    V = 0.5 * (0.5 * X**2 + 0.5 * Y**2)
    n = np.clip((mu - V) / g, 0.0, None)
    n = n * (1.0 / (np.sum(n) * dom.dx[0] * dom.dx[1] + 1e-12))
    dens = ys[-1][..., 0] ** 2 + ys[-1][..., 1] ** 2
    np.testing.assert_allclose(n, dens, rtol=1e-3, atol=1e-3)


def _exact_rhs(kind, N, L, kappa=1e-2):
    import sympy as sp

    x, y = sp.symbols("x y", real=True)
    u = sp.sin(2 * x) * sp.cos(3 * y)  # u_star at t=0 (reference tests/test_rhs_convergence.py:16,49)
    mu = u**3 - u - kappa * (sp.diff(u, x, 2) + sp.diff(u, y, 2))
    M = 1 + u**2
    if kind == "ac":
        f = -M * mu  # symbolic/allen_cahn_sym.py
    else:
        f = sp.diff(M * sp.diff(mu, x), x) + sp.diff(M * sp.diff(mu, y), y)  # symbolic/cahn_hilliard_sym.py
    dom = O.Domain((N, N), ((-L / 2, L / 2), (-L / 2, L / 2)))
    X, Y = dom.mesh()
    return dom, sp.lambdify((x, y), u, "numpy")(X, Y), sp.lambdify((x, y), f, "numpy")(X, Y)


@pytest.mark.parametrize("kind", ["ac", "ch"])
def test_rhs_fd_second_order(kind):
    """reference tests/test_rhs_convergence.py:14-44 / :47-77: fitted order 2 (rtol 0.1)."""
    errs, dxs = [], []
    for N in [32, 64, 128, 256, 512]:
        dom, u, f_exact = _exact_rhs(kind, N, 2 * np.pi)
        if kind == "ac":
            eq = O.AllenCahn2DPeriodic(dom, 1e-2, O.mu_double_well, lambda c: 1 + c**2, "fd", np.float64)
        else:
            eq = O.CahnHilliardPeriodic(dom, 1e-2, O.mu_double_well, lambda c: 1 + c**2, "fd", np.float64)
        f = eq.rhs(u, 0.0)
        errs.append(np.sqrt(np.sum((f - f_exact) ** 2)) / np.sqrt(np.sum(f_exact**2)))
        dxs.append(dom.dx[0])
    slope = np.polyfit(np.log(dxs), np.log(errs), 1)[0]
    np.testing.assert_allclose(slope, 2.0, rtol=0.1)


@pytest.mark.parametrize("kind", ["ac", "ch"])
def test_rhs_fourier_matches_exact(kind):
    """spectral RHS (cahn_hilliard.py:82-87, allen_cahn.py:74-79) is exact for a band-limited field."""
    dom, u, f_exact = _exact_rhs(kind, 64, 2 * np.pi)
    if kind == "ac":
        eq = O.AllenCahn2DPeriodic(dom, 1e-2, O.mu_double_well, lambda c: 1 + c**2, "fourier", np.float64)
    else:
        eq = O.CahnHilliardPeriodic(dom, 1e-2, O.mu_double_well, lambda c: 1 + c**2, "fourier", np.float64)
    f = eq.rhs(u, 0.0)
    assert np.linalg.norm(f - f_exact) / np.linalg.norm(f_exact) < 1e-10


def test_legendre_matches_legval():
    """reference tests/test_functions.py:22-61 (rtol=1e-5, atol=1e-7)."""
    from numpy.polynomial.legendre import legval

    rng = np.random.default_rng(0)
    for deg in [0, 1, 2, 5, 9]:
        p = rng.normal(size=deg + 1)
        x = np.linspace(-1, 1, 101)
        np.testing.assert_allclose(O.legendre_expansion(p, x), legval(x, p), rtol=1e-5, atol=1e-7)
        c = np.linspace(0.01, 0.99, 57)
        np.testing.assert_allclose(O.D_legendre(p, c), np.exp(legval(2 * c - 1, p)), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(
            O.mu_legendre(p, c, O.prior_log), legval(2 * c - 1, p) + np.log(c / (1 - c)), rtol=1e-5, atol=1e-7
        )


def test_advection_diffusion_fixture():
    """notebooks/reference.npy (run_advection_diffusion.ipynb cells 1-7, t=5): pins the form
    du/dt = -div(v u) + D lap u with spectral derivatives.  The jax PRNG initial
    noise cannot be reproduced; it has decayed by t=5, only its mean survives."""
    ref = np.load(os.path.join(GOLD, "ref_advection_diffusion_64.npy"))
    N = 64
    L = 0.02 * N
    dom = O.Domain((N, N), ((-L / 2, L / 2), (-L / 2, L / 2)))
    eq = O.AdvectionDiffusion2D(dom, O.gaussian_velocity([0.1, 0.01], (0.4, 0.4)), 0.1, np.float64)
    y = np.full((N, N), float(ref.mean()))
    t, dt = 0.0, 2e-3
    for _ in range(2500):
        y = O.sifs_step(eq.rhs, y, t, t + dt, 1.0, eq.fourier_symbol)
        t += dt
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-4


def test_constant_step_schedule():
    """diffrax loop restatement: 0.05 / 1e-4 in float32 (pde_env.py:293-303 with the
    notebook's step_dt / numeric_dt) visits ~500 steps and ends exactly on t1."""
    ts = O.constant_step_schedule(0.0, 0.05, 1e-4, np.float32)
    assert ts[-1] == np.float32(0.05) and ts[0] == 0
    assert abs(len(ts) - 1 - 500) <= 1
    assert np.all(np.diff(ts) > 0)
    ts64 = O.constant_step_schedule(0.0, 10.0, 5e-5, np.float64)
    assert abs(len(ts64) - 1 - 200000) <= 1 and ts64[-1] == 10.0


def test_saveat_interpolation_endpoints():
    y0 = np.zeros((4,), np.float64)
    ys = O.integrate(lambda y, a, b: y + (b - a), y0, 0.0, 1.0, 0.3, save_ts=[0.0, 0.45, 1.0])
    np.testing.assert_allclose(ys[:, 0], [0.0, 0.45, 1.0], atol=1e-12)


def test_torch_gradient_oracle_matches_numpy_oracle_and_finite_differences():
    """The torch float64 twin used as the gradient oracle (stand-in for jax.grad) reproduces the
    NumPy oracle's forward rollout and its autograd gradients agree with central differences."""
    import torch

    from oracle import ad_torch_oracle as TO

    n, h, D = 32, 0.04, 0.1
    box = ((-n * h / 2, n * h / 2),) * 2
    rng = np.random.default_rng(0)
    y0 = 0.5 + 0.01 * rng.normal(size=(1, n, n))
    ctrl = np.array([[[0.1, -0.2, 0.1, 0.02], [-0.1, 0.15, 0.15, 0.03]]])
    dts = np.full(6, 1e-3)
    dom = O.Domain((n, n), box)
    y = y0[0]
    for k, dt in enumerate(dts):
        cx, cy, p0, p1 = ctrl[0, min(k // 3, 1)]
        eq = O.AdvectionDiffusion2D(dom, O.gaussian_velocity((p0, p1), (cx, cy)), D, np.float64)
        y = O.sifs_step(eq.rhs, y, 0.0, dt, 1.0, eq.fourier_symbol)
    yt = torch.from_numpy(y0).requires_grad_(True)
    ct = torch.from_numpy(ctrl).requires_grad_(True)
    out = TO.rollout(yt, ct, dts, (n, n), box, D, 1.0, hold=3)
    assert np.allclose(out.detach().numpy()[0], y, rtol=1e-12, atol=1e-14)
    loss = (out**2).mean()
    loss.backward()

    def f(c):
        with torch.no_grad():
            return float((TO.rollout(torch.from_numpy(y0), torch.from_numpy(c), dts, (n, n), box, D, 1.0, hold=3) ** 2).mean())

    for idx in [(0, 0, 0), (0, 1, 1), (0, 0, 2), (0, 1, 3)]:
        eps = 1e-6
        cp, cm = ctrl.copy(), ctrl.copy()
        cp[idx] += eps
        cm[idx] -= eps
        fd = (f(cp) - f(cm)) / (2 * eps)
        assert abs(fd - float(ct.grad[idx])) <= 1e-5 * max(abs(fd), 1e-12) + 1e-14


def test_torch_phasefield_gradient_oracle_matches_numpy_oracle():
    """oracle/ch_torch_oracle.py (the float64 gradient oracle of the phase-field adjoint tests)
    reproduces the NumPy oracle's Cahn-Hilliard and Allen-Cahn rollouts, Legendre closures included."""
    import torch

    from oracle import ch_torch_oracle as TO

    n, h = 32, 0.01
    box = ((0.0, n * h), (0.0, n * h))
    u = np.clip(0.5 + 0.05 * np.random.default_rng(0).normal(size=(n, n)), 0.1, 0.9)
    pm, pd = [0.1, 2.5, -0.3, 0.8], [-1.0, 0.3, -0.2]
    dom = O.Domain((n, n), box)
    for kind, A in (("ch", 0.5), ("ac", 1.0)):
        if kind == "ch":
            eq = O.CahnHilliardPeriodic(dom, 0.002, lambda c: O.mu_legendre(pm, c, O.prior_log), lambda c: O.D_legendre(pd, c), "fd", np.float64)
        else:
            eq = O.AllenCahn2DPeriodic(dom, 0.002, lambda c: O.mu_legendre(pm, c, O.prior_log), lambda c: O.D_legendre(pd, c), "fd", np.float64)
        y = u
        for _ in range(4):
            y = O.sifs_step(eq.rhs, y, 0.0, 1e-6, A, eq.fourier_symbol)
        tm, td = torch.tensor(pm, dtype=torch.float64), torch.tensor(pd, dtype=torch.float64)
        yt = TO.rollout(torch.from_numpy(u)[None], [1e-6] * 4, (n, n), box, 0.002, A, lambda c: TO.mu_legendre(tm, c, True),
                        lambda c: TO.D_legendre(td, c), kind)
        assert np.abs(yt[0].numpy() - y).max() <= 1e-13


def test_detect_vortices_oracle_counts_known_windings():
    """pde_opt/rl_utils.py:19-84 restated: four singly quantised vortices at cell centres inside the
    grid; a periodic field has zero total circulation, so the remaining charge sits on the boundary
    cells where the non-periodic test field jumps — the amplitude mask removes those."""
    N = 64
    centres = [(20.5, 20.5, 1), (40.5, 24.5, -1), (30.5, 44.5, 1), (44.5, 44.5, 1)]
    psi = O.vortex_test_field(N, centres)
    amp = 0.15  # corner-averaged density is ~0.4 at the four cores and < 0.07 on the boundary
    r = O.detect_vortices(psi, amp_thresh=amp, tol=0.5)
    w = r["winding"]
    for (ci, cj, q) in centres:
        assert w[int(ci), int(cj)] == q
    assert r["num_vortices"] == 4 and r["total_topological_charge"] == 2 and r["abs_charge_count"] == 4
    assert np.array_equal(np.sort(r["charges"]), np.array([-1, 1, 1, 1]))
    assert r["positions"].shape == (4, 2) and np.allclose(r["positions"] % 1.0, 0.5)
    # without the mask the periodic plaquette sum vanishes identically (Stokes on a torus)
    assert O.detect_vortices(psi)["total_topological_charge"] == 0


def test_smoothed_boundary_oracle_reduces_to_the_periodic_equations_for_a_trivial_geometry():
    """With psi == 1 (no boundary: |grad psi| = 0) the smoothed-boundary right-hand sides (cahn_hilliard.py:261-289,
    allen_cahn.py:142-159) must equal the periodic ones (cahn_hilliard.py:89-109, allen_cahn.py:81-84), which the
    reference's own known-answer tests pin: a consistency check of the restatement in oracle/pde_oracle.py."""
    n, h, kappa = 32, 0.01, 0.002
    dom = O.Domain((n, n), ((0.0, n * h),) * 2)
    rng = np.random.default_rng(3)
    u = np.clip(0.5 + 0.2 * rng.normal(size=(n, n)), 0.05, 0.95)
    psi = np.ones((n, n))
    side = np.zeros((n, n)); side[: n // 2] = 1
    f = lambda c: c * np.log(c) + (1 - c) * np.log(1 - c) + 3.0 * c * (1 - c) + 0.059
    mu = lambda c: O.mu_log(c, 3.0)
    D = lambda c: (1 - c) * c
    ch = O.CahnHilliardPeriodic(dom, kappa, mu, D, "fd", np.float64)
    got = O.sbm_rhs_ch(u, 0.3, psi, dom.dx, kappa, f, mu, D, lambda t: 1.0, lambda t: 0.7, side)
    assert np.allclose(got, ch.rhs_fd(u), rtol=1e-10, atol=1e-8 * np.abs(got).max())
    ac = O.AllenCahn2DPeriodic(dom, kappa, mu, D, "fd", np.float64)
    got = O.sbm_rhs_ac(u, 0.3, psi, dom.dx, kappa, f, mu, D, lambda t: 1.0, side)
    assert np.allclose(got, ac.rhs_fd(u), rtol=1e-10, atol=1e-8 * np.abs(got).max())
    # and the boundary terms switch on with the geometry: a non-trivial psi changes the answer
    x = np.linspace(-1, 1, n)
    psi2 = np.clip(0.5 * (1 + np.tanh((0.6 - np.abs(x)) / 0.1))[:, None] * np.ones((1, n)), 0.001, 1.0)
    assert not np.allclose(O.sbm_rhs_ac(u, 0.3, psi2, dom.dx, kappa, f, mu, D, lambda t: 1.0, side), ac.rhs_fd(u))
