"""Line-FFT engine and the paths built on it: fftn/ifftn vs numpy.fft, 3-D Cahn-Hilliard (rhs and
semi-implicit steps) and Strang split-step on 256x256 / rectangular grids vs the NumPy oracle."""
import numpy as np
import pytest

from oracle import pde_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.complex128), np.asarray(b, np.complex128)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("n", [8, 16, 32, 64, 128, 256, 512])
def test_fft_lines_every_length_contiguous_and_strided(n):
    from pde_opt_b200.linefft import fft_lines, pos_to_freq

    rng = np.random.default_rng(n)
    x = (rng.normal(size=(3, n, 40)) + 1j * rng.normal(size=(3, n, 40))).astype(np.complex64)
    xt = torch.from_numpy(x).cuda()
    # strided lines (axis 1) — ragged tile (40 lines per batch entry is not a multiple of the tile)
    got = fft_lines(xt, 1).cpu().numpy()
    want = np.fft.fft(x.astype(np.complex128), axis=1)[:, pos_to_freq(n), :]
    assert rel_l2(got, want) <= 2e-6
    back = fft_lines(torch.from_numpy(got).cuda(), 1, inverse=True, scale=1.0 / n).cpu().numpy()
    assert rel_l2(back, x) <= 2e-6
    # contiguous lines (last axis)
    xc = np.ascontiguousarray(np.swapaxes(x, 1, 2))
    got = fft_lines(torch.from_numpy(xc).cuda(), 2).cpu().numpy()
    want = np.fft.fft(xc.astype(np.complex128), axis=2)[:, :, pos_to_freq(n)]
    assert rel_l2(got, want) <= 2e-6
    # real input
    got = fft_lines(torch.from_numpy(np.ascontiguousarray(xc.real)).cuda(), 2).cpu().numpy()
    want = np.fft.fft(xc.real.astype(np.float64), axis=2)[:, :, pos_to_freq(n)]
    assert rel_l2(got, want) <= 2e-6


def test_fftn_ifftn_match_numpy():
    from pde_opt_b200.linefft import fftn, ifftn

    rng = np.random.default_rng(0)
    for shape in [(64, 32), (16, 32, 64), (256, 256)]:
        x = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)
        X = fftn(torch.from_numpy(x).cuda())
        assert rel_l2(X.cpu().numpy(), np.fft.fftn(x.astype(np.complex128))) <= 3e-6
        assert rel_l2(ifftn(X).cpu().numpy(), x) <= 3e-6


# ---- 3-D Cahn-Hilliard --------------------------------------------------------------------------
def _ch3d(points, h, mu_name="log"):
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard3DPeriodic
    from pde_opt_b200.functions import ConstantMobility, DegenerateMobility, DoubleWell, LogRegular

    hs = h if isinstance(h, tuple) else (h, h, h)
    box = tuple((0.0, n * hh) for n, hh in zip(points, hs))
    dom, odom = Domain(points, box, "dimensionless"), O.Domain(points, box)
    if mu_name == "logdeg":  # notebooks/optimize_nn_script.py:33-37 closures on a 3-D grid
        eq = CahnHilliard3DPeriodic(dom, 0.002, LogRegular(3.0), DegenerateMobility())
        oeq = O.CahnHilliardPeriodic(odom, 0.002, lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, "fd", np.float32)
    elif mu_name == "dwconst":
        eq = CahnHilliard3DPeriodic(dom, 0.002, DoubleWell(), ConstantMobility(0.7))
        oeq = O.CahnHilliardPeriodic(odom, 0.002, O.mu_double_well, lambda c: 0.7 * np.ones_like(c), "fd", np.float32)
    elif mu_name == "log":  # docs/notebooks/optimization_3D.ipynb cell 8: log potential, D = 0.15
        eq = CahnHilliard3DPeriodic(dom, 0.002, LogRegular(3.0), ConstantMobility(0.15))
        oeq = O.CahnHilliardPeriodic(odom, 0.002, lambda c: O.mu_log(c, 3.0), lambda c: 0.15 * np.ones_like(c), "fd", np.float32)
    else:
        eq = CahnHilliard3DPeriodic(dom, 0.002, DoubleWell(), DegenerateMobility())
        oeq = O.CahnHilliardPeriodic(odom, 0.002, O.mu_double_well, lambda c: (1 - c) * c, "fd", np.float32)
    return eq, oeq


def _u0(points, B, seed=0):
    return np.stack([np.clip(0.5 + 0.01 * np.random.default_rng(seed + i).normal(size=points), 0.01, 0.99) for i in range(B)]).astype(np.float32)


@pytest.mark.parametrize("points,mu_name,h", [((32, 32, 32), "log", 0.01), ((16, 32, 64), "dw", 0.01),
                                              # register-marching kernel (ny % 16 == 0, nz % 64 == 0): every closure
                                              # specialisation, several tiles per axis, anisotropic spacing
                                              ((16, 32, 128), "log", 0.01), ((8, 48, 64), "logdeg", (0.01, 0.012, 0.008)),
                                              ((24, 16, 192), "dwconst", (0.02, 0.01, 0.015)), ((64, 64, 64), "dw", 0.01),
                                              # 32 x 32 tile of the marching kernel (nz % 64 != 0)
                                              ((16, 64, 32), "logdeg", (0.01, 0.012, 0.008)), ((8, 32, 96), "dw", 0.01)])
def test_ch3d_rhs_matches_oracle(points, mu_name, h):
    eq, oeq = _ch3d(points, h, mu_name)
    u = _u0(points, 2)
    f = eq.rhs(torch.from_numpy(u).cuda()).cpu().numpy()
    o64 = O.CahnHilliardPeriodic(oeq.domain, 0.002, oeq.mu, oeq.D, "fd", np.float64)
    for b in range(2):
        assert rel_l2(f[b], o64.rhs(u[b].astype(np.float64))) <= 2e-4


@pytest.mark.parametrize("points,mu_name,K", [((32, 32, 32), "log", 1), ((32, 32, 32), "log", 16), ((16, 32, 64), "dw", 5), ((64, 64, 64), "log", 3)])
def test_ch3d_steps_match_oracle(points, mu_name, K):
    from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral

    eq, oeq = _ch3d(points, 0.01, mu_name)
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    B = 2 if points[0] < 64 else 1
    u = _u0(points, B, 5)
    times = O.constant_step_schedule(0.0, K * 1e-6, 1e-6, np.float32)
    got = solver.rollout(ODETerm(eq), times, torch.from_numpy(u).cuda()).cpu().numpy()
    for b in range(B):
        y = u[b]
        for a, bb in zip(times[:-1], times[1:]):
            y = O.sifs_step(oeq.rhs, y, a, bb, 0.5, oeq.fourier_symbol)
        assert rel_l2(got[b], y) <= 1e-5
        assert rel_l2(got[b] - u[b], y - u[b]) <= 2e-3
    # single (unbatched) state through solver.step
    y1, err, dense, st, res = solver.step(ODETerm(eq), times[0], times[1], torch.from_numpy(u[0]).cuda())
    assert err is None and res == 0 and tuple(y1.shape) == points
    assert rel_l2(y1.cpu().numpy(), O.sifs_step(oeq.rhs, u[0], times[0], times[1], 0.5, oeq.fourier_symbol)) <= 1e-5


# ---- Strang split-step on grids that do not fit one SM --------------------------------------------
def _gpe(points, L_, k, light, e):
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import GPE2DTSControl

    box = ((-L_ / 2, L_ / 2), (-L_ / 2, L_ / 2))
    return Domain(points, box, "dimensionless"), O.Domain(points, box), None


@pytest.mark.parametrize("points", [(256, 256), (64, 128)])
@pytest.mark.parametrize("time_scale", [-1j, 1.0])
@pytest.mark.parametrize("kinetic", [False, True])
def test_strang_lines_match_oracle(points, time_scale, kinetic):
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import GPE2DTSControl
    from pde_opt_b200.functions import GaussianLight
    from pde_opt_b200.solvers import ODETerm, StrangSplitting
    from tests.test_strang_gpu import tf_setup

    L_, k, x_s, t_s = tf_setup()
    light = GaussianLight(5.0, 1.5, -2.0, 3.0)
    box = ((-L_ / 2, L_ / 2), (-L_ / 2, L_ / 2))
    dom, odom = Domain(points, box, "dimensionless"), O.Domain(points, box)
    eq = GPE2DTSControl(dom, k, 0.1, light, trap_factor=1.0)
    oeq = O.GPE2DTSControl(odom, k, 0.1, lambda t, x, y: light(t, x, y), 1.0, np.float32, kinetic=kinetic)
    a_term = oeq.A_term if kinetic else eq.A_term
    solver = StrangSplitting(a_term, eq.dx, eq.fft, eq.ifft, time_scale)
    i, j = np.meshgrid(np.arange(points[0]), np.arange(points[1]), indexing="ij")
    y0 = []
    for s in range(3):
        rng = np.random.default_rng(s)
        psi = np.exp(-(((i - points[0] // 2) / (0.4 * points[0])) ** 2) - ((j - points[1] // 2) / (0.4 * points[1])) ** 2).astype(complex)
        psi = psi * np.exp(0.3j * rng.normal(size=points)) * (1 + 0.05 * rng.normal(size=points))
        psi = psi / np.sqrt(np.sum(np.abs(psi) ** 2) * odom.dx[0] ** 2)
        y0.append(np.stack([psi.real, psi.imag], -1).astype(np.float32))
    y0 = np.stack(y0)
    dt_ = 1e-5 / t_s
    times = O.constant_step_schedule(0.0, 6 * dt_, dt_, np.float32)
    got = solver.rollout(ODETerm(eq), times, torch.from_numpy(y0).cuda()).cpu().numpy()
    for b in range(3):
        y = y0[b]
        for a, bb in zip(times[:-1], times[1:]):
            y = O.strang_step(oeq.B_terms, y, a, bb, oeq.A_term, oeq.dx, time_scale)
        assert rel_l2(got[b], y) <= 2e-5, (points, time_scale, kinetic, b)


def test_gpe_ground_state_fixture_is_nearly_stationary():
    """pde_opt/data/ground_state.npy of the reference (256x256x2 float32; committed as
    tests/golden/gpe_ground_state_256.npy) through imaginary-time steps of the shipped equation
    (A_term == 0): GPU == oracle, and the density changes little (it is a relaxed state)."""
    import os

    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import GPE2DTSControl
    from pde_opt_b200.solvers import ODETerm, StrangSplitting
    from tests.test_strang_gpu import tf_setup

    path = os.path.join(os.path.dirname(__file__), "golden", "gpe_ground_state_256.npy")
    g = np.load(path).astype(np.float32)
    assert g.shape == (256, 256, 2)
    L_, k, x_s, t_s = tf_setup()
    box = ((-L_ / 2, L_ / 2), (-L_ / 2, L_ / 2))
    dom, odom = Domain((256, 256), box, "dimensionless"), O.Domain((256, 256), box)
    eq = GPE2DTSControl(dom, k, 0.0, lambda t, x, y: 0.0 * x, trap_factor=1.0)
    oeq = O.GPE2DTSControl(odom, k, 0.0, lambda t, x, y: 0.0 * x, 1.0, np.float32)
    g = g / np.sqrt(np.sum(g.astype(np.float64) ** 2) * odom.dx[0] ** 2).astype(np.float32)
    solver = StrangSplitting(eq.A_term, eq.dx, eq.fft, eq.ifft, -1j)
    dt_ = 1e-5 / t_s
    times = O.constant_step_schedule(0.0, 4 * dt_, dt_, np.float32)
    got = solver.rollout(ODETerm(eq), times, torch.from_numpy(g).cuda()).cpu().numpy()
    y = g
    for a, bb in zip(times[:-1], times[1:]):
        y = O.strang_step(oeq.B_terms, y, a, bb, oeq.A_term, oeq.dx, -1j)
    assert rel_l2(got, y) <= 2e-5


def test_slab_path_on_one_gpu_matches_whole_domain_step():
    """SlabCahnHilliard3D with world_size 1 (halo planes wrap on the rank, packed geometry with one
    chunk, all-to-all = copy) must reproduce pdeopt_ch3d_step and the oracle."""
    from pde_opt_b200.parallel import SlabCahnHilliard3D
    from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral

    points = (32, 16, 64)
    eq, oeq = _ch3d(points, 0.01, "log")
    u = _u0(points, 1, 3)[0]
    slab = SlabCahnHilliard3D(eq, 0.5, device="cuda")
    y = torch.from_numpy(u).cuda()
    for k in range(3):
        y = slab.step(y, 1e-6)
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    times = np.arange(4, dtype=np.float32) * np.float32(1e-6)
    whole = solver.rollout(ODETerm(eq), times, torch.from_numpy(u).cuda()).cpu().numpy()
    yo = u
    for a, bb in zip(times[:-1], times[1:]):
        yo = O.sifs_step(oeq.rhs, yo, a, bb, 0.5, oeq.fourier_symbol)
    assert rel_l2(y.cpu().numpy(), yo) <= 1e-5
    assert rel_l2(y.cpu().numpy() - u, whole - u) <= 1e-3


@pytest.mark.parametrize("n", [8, 32, 128, 512])
def test_r2c_and_c2r_lines(n):
    """real lines -> natural-order half spectra (two lines per complex transform) and back."""
    import ctypes

    from pde_opt_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(n)
    L = 70  # even, not a multiple of the tile
    x = rng.normal(size=(L, n)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    half = torch.empty((L, n // 2 + 1), dtype=torch.complex64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pdeopt_fft_lines_r2c(ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(half.data_ptr()), n, L, st))
    assert rel_l2(half.cpu().numpy(), np.fft.rfft(x.astype(np.float64), axis=1)) <= 2e-6
    y0 = torch.from_numpy(rng.normal(size=(L, n)).astype(np.float32)).cuda()
    y1 = torch.empty_like(y0)
    _lib.check(lib.pdeopt_fft_lines_c2r_update(ctypes.c_void_p(half.data_ptr()), n, L, ctypes.c_void_p(y0.data_ptr()), ctypes.c_void_p(y1.data_ptr()), 0.5 / n, st))
    assert rel_l2(y1.cpu().numpy(), y0.cpu().numpy() + 0.5 * x) <= 2e-6


def test_rhs_fourier_on_the_line_engine_3d_and_small_2d():
    """derivs='fourier' on grids without a fused kernel (3-D, 64x64 2-D): the reference expression
    (cahn_hilliard.py:82-87, :165-175; allen_cahn.py:74-79) on linefft.fftn / ifftn, and the unfused
    semi-implicit step on top of it, against the oracle."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AllenCahn2DPeriodic, CahnHilliard2DPeriodic, CahnHilliard3DPeriodic
    from pde_opt_b200.functions import DegenerateMobility, LogRegular, OnePlusSquare
    from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral

    # 3-D Cahn-Hilliard
    pts = (16, 32, 8)
    box = tuple((0.0, n * 0.01) for n in pts)
    eq = CahnHilliard3DPeriodic(Domain(pts, box, "d"), 0.002, LogRegular(3.0), DegenerateMobility(), derivs="fourier")
    oeq = O.CahnHilliardPeriodic(O.Domain(pts, box), 0.002, lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, "fourier", np.float32)
    o64 = O.CahnHilliardPeriodic(O.Domain(pts, box), 0.002, lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, "fourier", np.float64)
    u = np.clip(0.5 + 0.05 * np.random.default_rng(0).normal(size=(2,) + pts), 0.05, 0.95).astype(np.float32)
    f = eq.rhs(torch.from_numpy(u).cuda()).cpu().numpy()
    for b in range(2):
        assert rel_l2(f[b], o64.rhs(u[b].astype(np.float64))) <= 5e-5
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    times = np.arange(4, dtype=np.float32) * np.float32(1e-6)
    got = solver.rollout(ODETerm(eq), times, torch.from_numpy(u).cuda()).cpu().numpy()
    for b in range(2):
        y = u[b]
        for a, bb in zip(times[:-1], times[1:]):
            y = O.sifs_step(oeq.rhs, y, a, bb, 0.5, oeq.fourier_symbol)
        assert rel_l2(got[b], y) <= 1e-5
    # 2-D 64x64 (no fused fourier kernel): RHS only
    box2 = ((0.0, 0.64), (0.0, 0.64))
    u2 = np.clip(0.5 + 0.05 * np.random.default_rng(1).normal(size=(2, 64, 64)), 0.05, 0.95).astype(np.float32)
    for kind in ("ch", "ac"):
        if kind == "ch":
            e2 = CahnHilliard2DPeriodic(Domain((64, 64), box2, "d"), 0.002, LogRegular(3.0), DegenerateMobility(), derivs="fourier")
            o2 = O.CahnHilliardPeriodic(O.Domain((64, 64), box2), 0.002, lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, "fourier", np.float64)
        else:
            e2 = AllenCahn2DPeriodic(Domain((64, 64), box2, "d"), 0.002, LogRegular(3.0), OnePlusSquare(), derivs="fourier")
            o2 = O.AllenCahn2DPeriodic(O.Domain((64, 64), box2), 0.002, lambda c: O.mu_log(c, 3.0), lambda c: 1 + c**2, "fourier", np.float64)
        assert not e2.fused
        f2 = e2.rhs(torch.from_numpy(u2).cuda()).cpu().numpy()
        for b in range(2):
            assert rel_l2(f2[b], o2.rhs(u2[b].astype(np.float64))) <= 5e-5


def test_pde_model_solve_3d_saveat():
    """PDEModel.solve on CahnHilliard3DPeriodic (docs/notebooks/optimization_3D.ipynb: 32^3, log potential,
    D = 0.15): SaveAt(ts) semantics against the oracle's constant-step integrator."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard3DPeriodic
    from pde_opt_b200.functions import ConstantMobility, LogRegular
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    pts = (32, 32, 32)
    box = tuple((0.0, n * 0.01) for n in pts)
    model = PDEModel(CahnHilliard3DPeriodic, Domain(pts, box, "d"), SemiImplicitFourierSpectral)
    u = _u0(pts, 1, 9)[0]
    ts = np.asarray([0.0, 4e-6, 1.05e-5], dtype=np.float32)
    ys = model.solve({"kappa": 0.002, "mu": LogRegular(3.0), "D": ConstantMobility(0.15)}, torch.from_numpy(u).cuda(), ts,
                     {"A": 0.5}, dt0=1e-6)
    assert tuple(ys.shape) == (3,) + pts
    oeq = O.CahnHilliardPeriodic(O.Domain(pts, box), 0.002, lambda c: O.mu_log(c, 3.0), lambda c: 0.15 * np.ones_like(c), "fd", np.float32)
    want = O.integrate(lambda y, a, b: O.sifs_step(oeq.rhs, y, a, b, 0.5, oeq.fourier_symbol), u, ts[0], ts[-1], 1e-6, save_ts=ts)
    got = ys.cpu().numpy()
    for i in range(3):
        assert rel_l2(got[i], want[i]) <= 1e-5
    assert rel_l2(got[2] - u, want[2] - u) <= 2e-3
