"""CPU tests of the host-side mirror of the reference API and of the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import pde_oracle as O
from pde_opt_b200 import Domain, check_equation_solver_compatibility, prepare_solver_params
from pde_opt_b200 import _lib, functions, schedule
from pde_opt_b200.equations import AllenCahn2DPeriodic, CahnHilliard2DPeriodic
from pde_opt_b200.fused import fold_symbol
from pde_opt_b200.solvers import SemiImplicitFourierSpectral

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dom(n=128, h=0.01):
    return Domain((n, n), ((-n * h / 2, n * h / 2),) * 2, "dimensionless")


def test_domain_matches_oracle():
    d, o = dom(64, 0.02), O.Domain((64, 64), ((-0.64, 0.64),) * 2)
    assert d.dx == o.dx and d.L == o.L
    for a, b in zip(d.axes(), o.axes()):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(d.fft_mesh(), o.fft_mesh()):
        np.testing.assert_array_equal(a, b)
    assert d.mesh()[0][3, 5] == d.axes()[0][3] and d.mesh()[1][3, 5] == d.axes()[1][5]  # indexing="ij"


def test_symbols_match_oracle():
    d = dom()
    eq = CahnHilliard2DPeriodic(d, 0.002, functions.DoubleWell(), functions.ConstantMobility(1.0))
    oeq = O.CahnHilliardPeriodic(O.Domain((128, 128), ((-0.64, 0.64),) * 2), 0.002, O.mu_double_well, lambda c: c, "fd", np.float32)
    np.testing.assert_allclose(eq.fourier_symbol, oeq.fourier_symbol, rtol=2e-6)
    ac = AllenCahn2DPeriodic(d, 0.002, functions.DoubleWell(), functions.ConstantMobility(1.0))
    oac = O.AllenCahn2DPeriodic(O.Domain((128, 128), ((-0.64, 0.64),) * 2), 0.002, O.mu_double_well, lambda c: c, "fd", np.float32)
    np.testing.assert_allclose(ac.fourier_symbol, oac.fourier_symbol, rtol=2e-6)
    q = fold_symbol(eq.fourier_symbol, 0.5)
    assert q.shape == (65, 65) and q.dtype == np.float32
    np.testing.assert_allclose(q, 0.5 * eq.fourier_symbol.real[:65, :65], rtol=1e-6)


def test_fold_symbol_rejects_non_even_or_complex():
    s = np.ones((8, 8), np.complex64)
    s[1, 2] = 5.0
    with pytest.raises(ValueError):
        fold_symbol(s)
    s = np.ones((8, 8), np.complex64) * (1 + 1j)
    with pytest.raises(ValueError):
        fold_symbol(s)


def test_solver_equation_protocol():
    """pde_opt/utils.py:6-53: class-level hasattr check and by-name attribute injection."""
    check_equation_solver_compatibility(SemiImplicitFourierSpectral, CahnHilliard2DPeriodic)
    check_equation_solver_compatibility(SemiImplicitFourierSpectral, AllenCahn2DPeriodic)

    class NoSymbol:
        fft = None
        ifft = None

    with pytest.raises(ValueError, match="fourier_symbol"):
        check_equation_solver_compatibility(SemiImplicitFourierSpectral, NoSymbol)
    eq = CahnHilliard2DPeriodic(dom(), 0.002, lambda c: c**3 - c, lambda c: 0 * c + 1.0)
    params = prepare_solver_params(SemiImplicitFourierSpectral, {"A": 0.5}, eq)
    assert set(params) == {"A", "fourier_symbol", "fft", "ifft"}
    s = SemiImplicitFourierSpectral(**params)
    assert s.order(None) == 1 and s.init(None, 0, 1, None, None) is None
    assert SemiImplicitFourierSpectral.required_equation_attrs == ["fourier_symbol", "fft", "ifft"]
    with pytest.raises(ValueError):
        CahnHilliard2DPeriodic(dom(), 0.002, lambda c: c, lambda c: c, derivs="nope")


def test_closure_recognition():
    r = functions.recognize
    assert r(lambda c: c**3 - c, "mu").family == "double_well"
    m = r(lambda c: np.log(c / (1.0 - c)) + 2.5 * (1.0 - 2.0 * c), "mu")
    assert m.family == "log" and abs(m.coef[0] - 2.5) < 1e-9
    assert r(lambda c: np.ones_like(c), "mob").descriptor() == ("const", (1.0,))
    assert r(lambda c: 0.15 * np.ones_like(c), "mob").descriptor() == ("const", (0.15,))
    assert r(lambda c: (1 - c) * c, "mob").family == "degenerate"
    assert r(lambda c: 1 + c**2, "mob").family == "one_plus_sq"
    assert r(lambda c: np.sin(c), "mu") is None
    eq = CahnHilliard2DPeriodic(dom(), 0.002, lambda c: np.sin(c), lambda c: 0 * c + 1.0)
    assert not eq.fused
    with pytest.raises(NotImplementedError):
        eq.plan()
    # closures evaluate like the oracle's
    c = np.linspace(0.05, 0.95, 11)
    np.testing.assert_allclose(functions.LogRegular(3.0)(c), O.mu_log(c, 3.0))
    p = np.array([0.3, -0.2, 0.5, 0.1])
    np.testing.assert_allclose(functions.DiffusionLegendrePolynomials(p)(c), O.D_legendre(p, c))
    np.testing.assert_allclose(functions.ChemicalPotentialLegendrePolynomials(p, "log")(c), O.mu_legendre(p, c, O.prior_log))


@pytest.mark.parametrize("args", [(0.0, 0.05, 1e-4, np.float32), (0.0, 0.02, 1e-6, np.float32), (0.0, 1.0, 0.3, np.float64)])
def test_schedule_matches_oracle(args):
    a = schedule.constant_step_times(*args)
    b = O.constant_step_schedule(*args)
    np.testing.assert_array_equal(a, b)
    assert a[-1] == np.dtype(args[3]).type(args[1])
    assert (schedule.step_lengths(a) > 0).all()


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports every function include/pdeopt_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "pdeopt_b200.h")).read()
    declared = set(re.findall(r"\b(pdeopt_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pdeopt_plan_desc"}
    assert declared, "header parse failed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_lib.EXPORTS) == declared
    assert lib.pdeopt_abi_version() == 4


def test_plan_validation_without_gpu():
    lib = _lib.load()
    d = _lib.PlanDesc()
    d.kind, d.derivs, d.nx, d.ny, d.hx, d.hy, d.kappa = _lib.KIND_CH2D, _lib.DERIVS_FD, 128, 128, 0.01, 0.01, 0.002
    h = ctypes.c_void_p()
    assert lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.OK
    assert lib.pdeopt_table_len(h) == 65 * 65
    assert lib.pdeopt_plan_destroy(h) == _lib.OK
    d.nx = 100
    assert lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_UNSUPPORTED
    assert b"128x128" in lib.pdeopt_last_error()
    d.nx, d.ny = 256, 256  # too large for the single-CTA kernels
    assert lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_UNSUPPORTED
    d.nx, d.ny = 64, 64
    assert lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.OK
    assert lib.pdeopt_table_len(h) == 33 * 33
    lib.pdeopt_plan_destroy(h)
    d.ny = 128
    d.nx, d.hx = 128, -1.0
    assert lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_INVALID
    d.hx, d.mu_family = 0.01, 9
    assert lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_INVALID
    assert lib.pdeopt_plan_create(None, ctypes.byref(h)) == _lib.ERR_INVALID


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpdeopt_b200.so")
    with pytest.raises(_lib.PdeOptError, match="no CPU fallback"):
        _lib.load()


def test_compute_entry_points_reject_bad_arguments_before_touching_the_gpu():
    """Error behaviour of the C ABI (include/pdeopt_b200.h: every function returns a status, never throws):
    argument validation happens before any CUDA call, so it can be checked on a machine without a GPU.
    Fake non-null pointers are never dereferenced on these paths."""
    lib = _lib.load()
    fake = ctypes.c_void_p(0x1000)
    dts = (ctypes.c_float * 600)(*([1e-6] * 600))
    d = _lib.PlanDesc()
    d.kind, d.derivs, d.nx, d.ny, d.hx, d.hy, d.kappa = _lib.KIND_CH2D, _lib.DERIVS_FD, 128, 128, 0.01, 0.01, 0.002
    h = ctypes.c_void_p()
    assert lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.OK
    step = lib.pdeopt_sifs_step_batched
    assert step(None, fake, fake, 1, 1, dts, fake, None, None, 0.0, 1.0, None, None) == _lib.ERR_INVALID
    assert step(h, None, fake, 1, 1, dts, fake, None, None, 0.0, 1.0, None, None) == _lib.ERR_INVALID
    assert step(h, fake, fake, 0, 1, dts, fake, None, None, 0.0, 1.0, None, None) == _lib.ERR_INVALID  # empty batch
    assert step(h, fake, fake, 1, 0, dts, fake, None, None, 0.0, 1.0, None, None) == _lib.ERR_INVALID
    assert step(h, fake, fake, 1, 513, dts, fake, None, None, 0.0, 1.0, None, None) == _lib.ERR_INVALID  # > PDEOPT_MAX_FUSED_STEPS
    assert lib.pdeopt_last_error() != b""
    assert lib.pdeopt_rhs_batched(h, None, fake, 1, None, None) == _lib.ERR_INVALID
    lib.pdeopt_plan_destroy(h)

    g = _lib.GpeDesc()
    g.nx, g.ny, g.hx, g.hy = 64, 64, 0.1, 0.1
    assert lib.pdeopt_strang_step_batched(ctypes.byref(g), fake, fake, 1, 1, dts, None, 0.0, -1.0, None, None) == _lib.ERR_UNSUPPORTED
    g.nx, g.ny = 128, 128
    assert lib.pdeopt_strang_step_batched(ctypes.byref(g), fake, fake, 0, 1, dts, None, 0.0, -1.0, None, None) == _lib.ERR_INVALID
    assert lib.pdeopt_strang_step_batched(ctypes.byref(g), None, fake, 1, 1, dts, None, 0.0, -1.0, None, None) == _lib.ERR_INVALID

    assert lib.pdeopt_fft_pos_to_freq(100, 0) == -1 and lib.pdeopt_fft_pos_to_freq(64, 64) == -1
    geom = _lib.LineGeom()
    assert lib.pdeopt_fft_lines(fake, fake, 100, ctypes.byref(geom), ctypes.byref(geom), 0, 0, 1.0, None) == _lib.ERR_UNSUPPORTED

    c = _lib.Ch3dDesc()
    c.nx, c.ny, c.nz, c.hx, c.hy, c.hz, c.kappa = 48, 32, 32, 0.01, 0.01, 0.01, 0.002
    assert lib.pdeopt_ch3d_step(ctypes.byref(c), fake, fake, 1, 1, dts, fake, fake, None) == _lib.ERR_UNSUPPORTED  # 48 is not 2^k
    c.nx = 32
    assert lib.pdeopt_ch3d_step(ctypes.byref(c), fake, fake, 1, 0, dts, fake, fake, None) == _lib.ERR_INVALID
    assert lib.pdeopt_ch3d_rhs(ctypes.byref(c), None, None, None, fake, fake, 1, None) == _lib.ERR_INVALID

    assert lib.pdeopt_gpe_detect_vortices(None, 1, 64, 64, 0.0, 0.5, None, fake, None) == _lib.ERR_INVALID
    assert lib.pdeopt_gpe_detect_vortices(fake, 0, 64, 64, 0.0, 0.5, None, fake, None) == _lib.ERR_INVALID


def test_round2_entry_points_reject_bad_arguments_before_touching_the_gpu():
    """Same contract for the entry points added in round 2 (rollouts that keep states, fused adjoint, tangents, given-mu
    right-hand side and adjoint, smoothed-boundary right-hand side, slab pushes)."""
    lib = _lib.load()
    fake = ctypes.c_void_p(0x1000)
    dts = (ctypes.c_float * 8)(*([1e-6] * 8))
    d = _lib.PlanDesc()
    d.kind, d.derivs, d.nx, d.ny, d.hx, d.hy, d.kappa = _lib.KIND_CH2D, _lib.DERIVS_FD, 128, 128, 0.01, 0.01, 0.002
    h = ctypes.c_void_p()
    assert lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.OK
    fwd = lib.pdeopt_sifs_rollout_fwd
    assert fwd(h, fake, fake, 1, 4, dts, fake, None, 1, None) == _lib.ERR_INVALID       # no trajectory buffer
    assert fwd(h, fake, fake, 1, 4, dts, fake, fake, 0, None) == _lib.ERR_INVALID        # save_every < 1
    assert fwd(h, fake, fake, 1, 4, dts, fake, fake, 513, None) == _lib.ERR_INVALID      # > PDEOPT_MAX_FUSED_STEPS
    assert fwd(h, fake, fake, 0, 4, dts, fake, fake, 1, None) == _lib.ERR_INVALID
    bwd = lib.pdeopt_sifs_rollout_bwd
    assert bwd(h, None, fake, fake, 1, 4, dts, fake, None, fake, fake, None) == _lib.ERR_INVALID
    assert bwd(h, fake, fake, fake, 1, 0, dts, fake, None, fake, fake, None) == _lib.ERR_INVALID
    tan = lib.pdeopt_phasefield_tangent_steps
    assert tan(h, fake, fake, 1, 0, 4, dts, fake, fake, fake, fake, None) == _lib.ERR_INVALID  # ndir < 1
    assert tan(h, fake, None, 1, 2, 4, dts, fake, fake, fake, fake, None) == _lib.ERR_INVALID
    assert lib.pdeopt_phasefield_tangent_work_floats(h, 3, 2) == (2 + 3 * 2) * 3 * 128 * 128
    assert lib.pdeopt_phasefield_tangent_work_floats(h, 0, 2) == 0
    assert lib.pdeopt_rhs_given_mu_batched(h, fake, None, None, fake, 1, fake, None) == _lib.ERR_INVALID
    assert lib.pdeopt_rhs_given_mu_batched(h, fake, fake, None, fake, 0, fake, None) == _lib.ERR_INVALID
    assert lib.pdeopt_phasefield_adjoint_given_mu(h, fake, fake, None, fake, fake, fake, None, 1, 1e-6, fake, fake, None) == _lib.ERR_INVALID
    lib.pdeopt_plan_destroy(h)
    d.derivs = _lib.DERIVS_FOURIER
    assert lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.OK
    assert tan(h, fake, fake, 1, 1, 4, dts, fake, fake, fake, fake, None) == _lib.ERR_UNSUPPORTED  # tangents: derivs='fd' only
    assert bwd(h, fake, fake, fake, 1, 4, dts, fake, None, fake, fake, None) == _lib.ERR_UNSUPPORTED
    lib.pdeopt_plan_destroy(h)

    sd = _lib.SbmDesc(kind=_lib.KIND_CH2D, nx=64, ny=64, hx=0.01, hy=0.01, kappa=0.002)
    sbm = lib.pdeopt_sbm_rhs_batched
    assert sbm(ctypes.byref(sd), fake, fake, fake, fake, fake, fake, fake, 0.0, 0.0, 0.0, None, fake, 1, None) == _lib.ERR_INVALID  # CH needs work
    assert sbm(ctypes.byref(sd), fake, fake, fake, fake, None, fake, fake, 0.0, 0.0, 0.0, fake, fake, 1, None) == _lib.ERR_INVALID
    sd.kind = 3
    assert sbm(ctypes.byref(sd), fake, fake, fake, fake, fake, fake, fake, 0.0, 0.0, 0.0, fake, fake, 1, None) == _lib.ERR_INVALID

    peers = (ctypes.c_void_p * 2)(fake, fake)
    assert lib.pdeopt_push_blocks_to_peers(fake, peers, 2, 24, 0, 0, None) == _lib.ERR_INVALID   # not a multiple of 16 bytes
    assert lib.pdeopt_push_blocks_to_peers(fake, peers, 9, 32, 0, 0, None) == _lib.ERR_INVALID   # > 8 peers
    assert lib.pdeopt_push_rows_to_peers(fake, peers, 2, 64, 0, 0, 32, 64, 0, None) == _lib.ERR_INVALID  # n_rows < 1
    assert lib.pdeopt_push_rows_to_peers(None, peers, 2, 64, 0, 1, 32, 64, 0, None) == _lib.ERR_INVALID


def test_pid_controller_matches_oracle_restatement_on_a_scalar_ode():
    """pde_opt_b200.stepsize.PIDController against oracle.integrate_adaptive on y' = -50 y with an
    implicit/explicit Euler pair (the error estimate of solvers.py:61-65): same accept / reject
    sequence, same step sizes (CPU tensors; the controller itself needs no GPU)."""
    import torch

    from pde_opt_b200.stepsize import PIDController

    lam = 50.0

    def step_err(y, a, b):
        dt = np.float32(b - a)
        y1 = (y / (1 + lam * dt)).astype(np.float32)       # implicit Euler
        return y1, y1 - (y + dt * (-lam * y))                 # minus explicit Euler

    for pid in ({}, {"pcoeff": 0.3, "icoeff": 0.4}, {"pcoeff": 0.1, "icoeff": 0.3, "dcoeff": 0.05}):
        y0 = np.array([1.0, 0.5], np.float32)
        _, acc, rej = O.integrate_adaptive(step_err, y0, 0.0, 0.2, 1e-4, 1e-3, 1e-6, **pid)
        ctl = PIDController(1e-3, 1e-6, **pid)
        t, t1, dt, state, y = np.float32(0), np.float32(0.2), 1e-4, ctl.init_state(), y0
        a2 = r2 = 0
        while t < t1:
            tn = np.float32(t + np.float32(dt))
            if tn > t1 - np.float32(1e-6):
                tn = t1
            y1, e = step_err(y, t, tn)
            err = ctl.scaled_error(torch.from_numpy(y), torch.from_numpy(y1), torch.from_numpy(e))
            keep, dt, state = ctl.adapt(float(tn - t), err, 1, state)
            if keep:
                y, t, a2 = y1, tn, a2 + 1
            else:
                r2 += 1
        assert (a2, r2) == (acc, rej) and acc > 5
        np.testing.assert_allclose(y, np.exp(-lam * 0.2) * y0, atol=2e-3)


def test_levenberg_marquardt_iteration_on_analytic_problems():
    """pde_opt_b200.least_squares.lm_iterate (the host loop behind PDEModel.train(method="least_squares"), optimistix's
    LevenbergMarquardt restated) on problems with known answers: a linear least-squares problem is solved to machine
    precision; Rosenbrock in residual form converges to (1, 1); the loss never increases along accepted steps."""
    import torch

    from pde_opt_b200.least_squares import lm_iterate

    rng = np.random.default_rng(0)
    A = torch.from_numpy(rng.normal(size=(20, 3)))
    b = torch.from_numpy(rng.normal(size=20))

    def lin(theta):
        r = A @ theta - b
        return 0.5 * float(r @ r), A.T @ A, A.T @ r

    theta, hist = lm_iterate(lin, torch.zeros(3, dtype=torch.float64), max_steps=200)
    want = torch.linalg.lstsq(A, b[:, None]).solution[:, 0]
    assert torch.allclose(theta, want, atol=1e-7) and all(h1 <= h0 for h0, h1 in zip(hist, hist[1:]))

    def rosen(theta):
        x, y = float(theta[0]), float(theta[1])
        r = torch.tensor([10.0 * (y - x * x), 1.0 - x], dtype=torch.float64)
        J = torch.tensor([[-20.0 * x, 10.0], [-1.0, 0.0]], dtype=torch.float64)
        return 0.5 * float(r @ r), J.T @ J, J.T @ r

    theta, hist = lm_iterate(rosen, torch.tensor([-1.2, 1.0], dtype=torch.float64), max_steps=200)
    assert torch.allclose(theta, torch.ones(2, dtype=torch.float64), atol=1e-6), theta
    assert hist[-1] < 1e-12 and all(h1 <= h0 for h0, h1 in zip(hist, hist[1:]))

    # a non-finite trial point is rejected and the step size shrinks until a finite one is found
    def wall(theta):
        x = float(theta[0])
        if x > 2.0:
            return float("nan"), torch.eye(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)
        r = torch.tensor([x - 1.9], dtype=torch.float64)
        return 0.5 * float(r @ r), torch.ones((1, 1), dtype=torch.float64) * 1e-6, torch.tensor([(x - 1.9) * 1e-3], dtype=torch.float64)

    theta, hist = lm_iterate(wall, torch.tensor([0.0], dtype=torch.float64), max_steps=50)
    assert np.isfinite(hist[-1]) and float(theta[0]) <= 2.0
