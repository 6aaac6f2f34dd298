"""Parity cases the round-1 review found missing ("What's weak" 1-5), all through the C ABI:

* phase-separated Cahn-Hilliard states (c in [0.03, 0.97] with interfaces), where 1/(c(1-c)) and the fast
  log2 are hardest and where the dynamics of the reference's notebooks actually live;
* an environment's result does not depend on its batch neighbours - including a NaN neighbour - and the
  per-environment non-finite flags mark exactly the broken ones;
* `lights(t, x, y)` callables that vanish at t = 0 (unfused Strang path with the caller-evaluated field);
* `eq.fft / eq.ifft` are working callables (jnp.fft.fftn / ifftn convention);
* derivs='fourier' through the host-buffer entry point with chunks on concurrent streams."""
import numpy as np
import pytest

from oracle import pde_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N, H, KAPPA = 128, 0.01, 0.002


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))


def droplets(seed, lo, hi, n=N, ndrops=5):
    """tanh-interface droplets between the bulk values lo and hi (interface width sqrt(2 kappa) ~ 6 cells)."""
    rng = np.random.default_rng(seed)
    L = n * H
    x = (np.arange(n) + 0.5) * H
    X, Y = np.meshgrid(x, x, indexing="ij")
    phi = -np.ones((n, n))
    w = np.sqrt(2 * KAPPA)
    for _ in range(ndrops):
        cx, cy, r = rng.uniform(0, L, 2).tolist() + [rng.uniform(0.12, 0.25) * L]
        dx = np.minimum(np.abs(X - cx), L - np.abs(X - cx))
        dy = np.minimum(np.abs(Y - cy), L - np.abs(Y - cy))
        phi = np.maximum(phi, np.tanh((r - np.sqrt(dx * dx + dy * dy)) / w))
    c = 0.5 * (lo + hi) + 0.5 * (hi - lo) * phi + 0.002 * rng.normal(size=(n, n))
    return np.clip(c, min(lo, hi) - 0.01, max(lo, hi) + 0.01).astype(np.float32)


LEG_MU = (0.1, 2.5, -0.3, 0.8, 0.05)
LEG_D = (-0.5, 0.3, -0.2)
SEPARATED = {
    "log_degenerate": (("log", (3.0,)), ("degenerate", ()), lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, 0.03, 0.97),
    "log_const": (("log", (3.0,)), ("const", (1.0,)), lambda c: O.mu_log(c, 3.0), lambda c: np.ones_like(c), 0.03, 0.97),
    "legendre_logprior_exp": (
        ("legendre_logprior", LEG_MU), ("legendre_exp", LEG_D),
        lambda c: O.mu_legendre(np.asarray(LEG_MU, np.float32), c, O.prior_log),
        lambda c: O.D_legendre(np.asarray(LEG_D, np.float32), c), 0.04, 0.96,
    ),
    "dw_const": (("double_well", ()), ("const", (1.0,)), O.mu_double_well, lambda c: np.ones_like(c), -0.97, 0.97),
}


def _eq(mu_f, D_f, dtype=np.float32):
    dom = O.Domain((N, N), ((-N * H / 2, N * H / 2),) * 2)
    return O.CahnHilliardPeriodic(dom, KAPPA, mu_f, D_f, "fd", dtype)


@pytest.mark.parametrize("case", list(SEPARATED))
def test_phase_separated_states(case):
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    mu, mob, mu_f, D_f, lo, hi = SEPARATED[case]
    y0 = np.stack([droplets(s, lo, hi) for s in range(3)])
    plan = SifsPlan("ch2d", N, N, (-N * H / 2,) * 2, (H, H), KAPPA, mu, mob)
    eq, eq64 = _eq(mu_f, D_f), _eq(mu_f, D_f, np.float64)
    sym = torch.from_numpy(fold_symbol(eq.fourier_symbol, 0.5)).cuda()
    yd = torch.from_numpy(y0).cuda()
    # the right-hand side itself against the float64 oracle (no |y0| to hide behind)
    f = plan.rhs(yd).cpu().numpy()
    for b in range(3):
        want = eq64.rhs(y0[b].astype(np.float64))
        assert rel_l2(f[b], want) <= 2e-4, (case, b, rel_l2(f[b], want))
        # and no worse than the float32 NumPy oracle is itself
        assert rel_l2(f[b], want) <= 4 * rel_l2(eq.rhs(y0[b]), want) + 2e-5, (case, b)
    # (D = 1 with the log potential overshoots (0, 1) at the larger step: oracle and kernel both go NaN)
    big = 2e-5 if case in ("log_degenerate", "legendre_logprior_exp") else 2e-6
    for dt, K in ((1e-6, 1), (1e-6, 16), (big, 8)):
        got = plan.step(yd, [dt] * K, sym).cpu().numpy()
        for b in range(3):
            y, t = y0[b], np.float32(0)
            for _ in range(K):
                y = O.sifs_step(eq.rhs, y, t, t + np.float32(dt), 0.5, eq.fourier_symbol)
                t = t + np.float32(dt)
            assert rel_l2(got[b], y) <= 1e-5, (case, dt, K, b)
            assert rel_l2(got[b] - y0[b], y - y0[b]) <= 2e-3, (case, dt, K, b, rel_l2(got[b] - y0[b], y - y0[b]))
            assert got[b].min() > min(lo, hi) - 0.05 and got[b].max() < max(lo, hi) + 0.05


def test_phase_separated_long_rollout():
    """Spinodal decomposition run INTO the phase-separated regime by the kernel itself (4000 steps of
    dt = 2e-5 from noise) tracks the oracle to the 1000-step tolerance over its last 1000 steps."""
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    mu, mob, mu_f, D_f, _, _ = SEPARATED["log_degenerate"]
    plan = SifsPlan("ch2d", N, N, (-N * H / 2,) * 2, (H, H), KAPPA, mu, mob)
    eq = _eq(mu_f, D_f)
    sym = torch.from_numpy(fold_symbol(eq.fourier_symbol, 0.5)).cuda()
    y0 = np.clip(0.5 + 0.05 * np.random.default_rng(7).normal(size=(1, N, N)), 0.02, 0.98).astype(np.float32)
    dt = 2e-5
    y = torch.from_numpy(y0).cuda()
    for _ in range(6):
        y = plan.step(y, [dt] * 500, sym)
    mid = y.cpu().numpy()[0]
    assert mid.min() < 0.2 and mid.max() > 0.8, "the run did not phase-separate"
    got = plan.step(plan.step(y, [dt] * 500, sym), [dt] * 500, sym).cpu().numpy()[0]
    ref, t = mid, np.float32(0)
    for _ in range(1000):
        ref = O.sifs_step(eq.rhs, ref, t, t + np.float32(dt), 0.5, eq.fourier_symbol)
        t = t + np.float32(dt)
    assert rel_l2(got, ref) <= 1e-3
    assert rel_l2(got - mid, ref - mid) <= 2e-2


def test_environment_independent_of_its_neighbours_and_nonfinite_flags():
    """128x128 fd: one CTA per environment, no shared arithmetic.  The result of an environment is bit
    for bit the same whatever its batch neighbours are - also when a neighbour is NaN / Inf - and the
    flags mark exactly the broken environments."""
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    mu, mob, mu_f, D_f, lo, hi = SEPARATED["log_degenerate"]
    plan = SifsPlan("ch2d", N, N, (-N * H / 2,) * 2, (H, H), KAPPA, mu, mob)
    sym = torch.from_numpy(fold_symbol(_eq(mu_f, D_f).fourier_symbol, 0.5)).cuda()
    a, b, c = droplets(1, lo, hi), droplets(2, lo, hi), droplets(3, lo, hi)
    bad_nan = np.full((N, N), np.nan, np.float32)
    bad_inf = b.copy()
    bad_inf[17, 33] = np.inf
    bad_neg = b.copy()
    bad_neg[5, 5] = -0.5  # log of a negative concentration -> NaN produced by the step itself
    dts = [1e-6] * 16
    batches = {
        "plain": [a, b, c, a],
        "moved": [c, a, b],
        "nan": [bad_nan, a, bad_inf, c, bad_neg],
    }
    out, flags = {}, {}
    for k, v in batches.items():
        y = torch.from_numpy(np.stack(v)).cuda()
        fl = torch.full((len(v),), -1, dtype=torch.int32, device="cuda")
        out[k] = plan.step(y, dts, sym, nonfinite=fl).cpu().numpy()
        flags[k] = fl.cpu().numpy()
    np.testing.assert_array_equal(out["plain"][0], out["plain"][3])
    np.testing.assert_array_equal(out["plain"][0], out["moved"][1])
    np.testing.assert_array_equal(out["plain"][0], out["nan"][1])
    np.testing.assert_array_equal(out["plain"][2], out["nan"][3])
    np.testing.assert_array_equal(flags["plain"], [0, 0, 0, 0])
    np.testing.assert_array_equal(flags["nan"], [1, 0, 1, 0, 1])
    assert np.isfinite(out["nan"][1]).all() and np.isfinite(out["nan"][3]).all()


def test_nonfinite_flags_on_the_other_kernels():
    """Small grids (pair-packed kernels: the flags come from the streaming check) and GPE states."""
    import ctypes

    from pde_opt_b200 import _lib
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    n, h = 64, 0.01
    plan = SifsPlan("ch2d", n, n, (-n * h / 2,) * 2, (h, h), KAPPA, ("double_well", ()), ("const", (1.0,)))
    dom = O.Domain((n, n), ((-n * h / 2, n * h / 2),) * 2)
    eq = O.CahnHilliardPeriodic(dom, KAPPA, O.mu_double_well, lambda c: np.ones_like(c), "fd", np.float32)
    sym = torch.from_numpy(fold_symbol(eq.fourier_symbol, 0.5)).cuda()
    y0 = (0.01 * np.random.default_rng(0).normal(size=(4, n, n))).astype(np.float32)
    y0[2, 3, 3] = np.nan
    fl = torch.full((4,), -1, dtype=torch.int32, device="cuda")
    plan.step(torch.from_numpy(y0).cuda(), [1e-6] * 4, sym, nonfinite=fl)
    got = fl.cpu().numpy()
    assert got[2] == 1 and got[0] == 0 and got[1] == 0  # env 3 shares a CTA with env 2 in the small-grid kernel
    psi = torch.randn((3, 32, 32, 2), device="cuda")
    psi[1, 4, 4, 1] = float("inf")
    fl = torch.empty(3, dtype=torch.int32, device="cuda")
    lib = _lib.load()
    _lib.check(lib.pdeopt_nonfinite_flags(ctypes.c_void_p(psi.data_ptr()), 3, 32 * 32 * 2, ctypes.c_void_p(fl.data_ptr()),
                                          _lib.stream_ptr(psi)))
    np.testing.assert_array_equal(fl.cpu().numpy(), [0, 1, 0])


@pytest.mark.parametrize("n", [128, 64])
@pytest.mark.parametrize("kinetic", [False, True])
def test_time_dependent_lights_take_the_unfused_path(n, kinetic):
    """A light that is zero at t = 0 (lambda t, x, y: a t x) must not be mistaken for "no light"."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import GPE2DTSControl
    from pde_opt_b200.solvers import ODETerm, StrangSplitting

    L_ = 20.0
    box = ((-L_ / 2, L_ / 2),) * 2
    k = 300.0

    def lights(t, x, y):
        return 4.0e3 * t * x + 2.0e3 * t * t * y * y

    dom, odom = Domain((n, n), box, "dimensionless"), O.Domain((n, n), box)
    eq = GPE2DTSControl(dom, k, 0.1, lights, 1.0)
    assert not eq.fused
    oeq = O.GPE2DTSControl(odom, k, 0.1, lights, 1.0, np.float32, kinetic=kinetic)
    a_term = oeq.A_term if kinetic else eq.A_term
    solver = StrangSplitting(a_term, eq.dx, eq.fft, eq.ifft, 1.0)
    rng = np.random.default_rng(n)
    X, Y = odom.mesh()
    psi = np.exp(-(X**2 + Y**2) / 18.0) * np.exp(0.3j * rng.normal(size=(n, n)))
    psi = psi / np.sqrt(np.sum(np.abs(psi) ** 2) * odom.dx[0] ** 2)
    y0 = np.stack([psi.real, psi.imag], -1).astype(np.float32)
    dt_ = 2e-3
    times = O.constant_step_schedule(0.0, 6 * dt_, dt_, np.float32)
    got = solver.rollout(ODETerm(eq), times, torch.from_numpy(y0).cuda()).cpu().numpy()
    y = y0
    for a, b in zip(times[:-1], times[1:]):
        y = O.strang_step(oeq.B_terms, y, a, b, oeq.A_term, oeq.dx, 1.0)
    assert rel_l2(got, y) <= 5e-5
    # the light matters: dropping it (what the t = 0 probe used to do) is far outside the tolerance
    oeq0 = O.GPE2DTSControl(odom, k, 0.1, lambda t, x, y: 0.0 * x, 1.0, np.float32, kinetic=kinetic)
    y_no = y0
    for a, b in zip(times[:-1], times[1:]):
        y_no = O.strang_step(oeq0.B_terms, y_no, a, b, oeq0.A_term, oeq0.dx, 1.0)
    assert rel_l2(y_no, y) > 1e-2
    # B_terms evaluates the callable at the requested time
    bt = eq.B_terms(torch.from_numpy(y0).cuda(), float(times[3])).cpu().numpy()
    np.testing.assert_allclose(bt, oeq.B_terms(y0, times[3]), rtol=2e-5, atol=2e-4)


def test_equation_fft_attributes_are_callables():
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard2DPeriodic, CahnHilliard3DPeriodic
    from pde_opt_b200.functions import ConstantMobility, DoubleWell

    dom = Domain((64, 32), ((0.0, 0.64), (0.0, 0.32)), "dimensionless")
    eq = CahnHilliard2DPeriodic(dom, KAPPA, DoubleWell(), ConstantMobility(1.0))
    x = torch.randn((3, 64, 32), device="cuda")
    F = eq.fft(x)
    ref = torch.fft.fftn(x, dim=(1, 2))
    assert F.dtype == torch.complex64 and float((F - ref).abs().max() / ref.abs().max()) < 2e-6
    back = eq.ifft(F)
    assert float((back.real - x).abs().max()) < 1e-5 and float(back.imag.abs().max()) < 1e-5
    dom3 = Domain((16, 32, 8), ((0.0, 1.0),) * 3, "dimensionless")
    eq3 = CahnHilliard3DPeriodic(dom3, KAPPA, DoubleWell(), ConstantMobility(1.0))
    x3 = torch.randn((16, 32, 8), device="cuda")
    assert float((eq3.fft(x3) - torch.fft.fftn(x3)).abs().max()) < 1e-3


def test_fourier_plan_through_the_host_entry_point():
    """derivs='fourier' keeps one scratch line per environment; the host-buffer entry point launches
    chunks of one batch on concurrent streams: results must equal the device-pointer call."""
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    B = 640
    plan = SifsPlan("ch2d", N, N, (-N * H / 2,) * 2, (H, H), KAPPA, ("log", (3.0,)), ("degenerate", ()), "fourier")
    sym = fold_symbol(_eq(O.mu_double_well, lambda c: c).fourier_symbol, 0.5)
    rng = np.random.default_rng(5)
    y0 = np.clip(0.5 + 0.05 * rng.normal(size=(B, N, N)), 0.05, 0.95).astype(np.float32)
    dts = [1e-6, 1e-6]
    y1h = plan.step_host(y0, dts, sym)
    y1d = plan.step(torch.from_numpy(y0).cuda(), dts, torch.from_numpy(sym).cuda()).cpu().numpy()
    np.testing.assert_array_equal(y1h, y1d)
