"""GPU parity of the fused 128x128 SIFS kernels against the NumPy oracle (through the C ABI)."""
import numpy as np
import pytest

from oracle import pde_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N = 128
H = 0.01
KAPPA = 0.002


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))


def make_ic(B, seed0=0, centre=0.5):
    return np.stack(
        [np.clip(centre + 0.01 * np.random.default_rng(seed0 + i).normal(size=(N, N)), 0.0, 1.0) for i in range(B)]
    ).astype(np.float32)


def oracle_eq(kind, mu, D):
    dom = O.Domain((N, N), ((-N * H / 2, N * H / 2),) * 2)
    if kind == "ch2d":
        return O.CahnHilliardPeriodic(dom, KAPPA, mu, D, "fd", np.float32)
    return O.AllenCahn2DPeriodic(dom, KAPPA, mu, D, "fd", np.float32)


LEG_MU = (0.1, 2.5, -0.3, 0.8, 0.05)
LEG_D = (-0.5, 0.3, -0.2)

CASES = {
    "log_degenerate": (("log", (3.0,)), ("degenerate", ()), lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, 0.5),
    "log_const": (("log", (3.0,)), ("const", (1.0,)), lambda c: O.mu_log(c, 3.0), lambda c: np.ones_like(c), 0.5),
    "dw_const": (("double_well", ()), ("const", (1.0,)), O.mu_double_well, lambda c: np.ones_like(c), 0.0),
    "dw_1psq": (("double_well", ()), ("one_plus_sq", ()), O.mu_double_well, lambda c: 1 + c * c, 0.0),
    # runtime-switch kernel: Legendre closures (functions/legendre.py:37-74)
    "legendre_logprior_exp": (
        ("legendre_logprior", LEG_MU), ("legendre_exp", LEG_D),
        lambda c: O.mu_legendre(np.asarray(LEG_MU, np.float32), c, O.prior_log),
        lambda c: O.D_legendre(np.asarray(LEG_D, np.float32), c), 0.5,
    ),
    "legendre_const": (
        ("legendre", LEG_MU), ("const", (0.15,)),
        lambda c: O.mu_legendre(np.asarray(LEG_MU, np.float32), c, None), lambda c: 0.15 * np.ones_like(c), 0.5,
    ),
    "log_1psq": (("log", (2.5,)), ("one_plus_sq", ()), lambda c: O.mu_log(c, 2.5), lambda c: 1 + c * c, 0.5),
}


def run_gpu(kind, mu, mob, y0, dts, A):
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    plan = SifsPlan(kind, N, N, (-N * H / 2, -N * H / 2), (H, H), KAPPA, mu, mob)
    eq = oracle_eq(kind, O.mu_double_well, lambda c: c)
    sym = torch.from_numpy(fold_symbol(eq.fourier_symbol, A)).cuda()
    yd = torch.from_numpy(y0).cuda()
    out = plan.step(yd, dts, sym)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("case", list(CASES))
def test_ch_one_step(case):
    mu, mob, mu_f, D_f, centre = CASES[case]
    y0 = make_ic(5, 0, centre)  # odd batch exercises the duplicated last env
    dt = 1e-6
    got = run_gpu("ch2d", mu, mob, y0, [dt], 0.5)
    eq = oracle_eq("ch2d", mu_f, D_f)
    for b in range(y0.shape[0]):
        ref = O.sifs_step(eq.rhs, y0[b], np.float32(0), np.float32(dt), 0.5, eq.fourier_symbol)
        assert rel_l2(got[b], ref) <= 1e-5, (case, b)
        # the increment itself (not masked by |y0|) must agree to fp32 FFT accuracy
        assert rel_l2(got[b] - y0[b], ref - y0[b]) <= 2e-3, (case, b)


def test_ch_k16_varying_step_lengths():
    mu, mob, mu_f, D_f, centre = CASES["log_degenerate"]
    y0 = make_ic(4, 10, centre)
    dts = [1e-6 * (1 + 1e-3 * np.sin(i)) for i in range(15)] + [5e-7]
    got = run_gpu("ch2d", mu, mob, y0, dts, 0.5)
    eq = oracle_eq("ch2d", mu_f, D_f)
    for b in range(4):
        y, t = y0[b], np.float32(0)
        for d in dts:
            y = O.sifs_step(eq.rhs, y, t, t + np.float32(d), 0.5, eq.fourier_symbol)
            t = t + np.float32(d)
        assert rel_l2(got[b], y) <= 1e-5


def test_ch_1000_steps():
    mu, mob, mu_f, D_f, centre = CASES["log_degenerate"]
    y0 = make_ic(2, 20, centre)
    dts = [1e-6] * 1000
    got = run_gpu("ch2d", mu, mob, y0, dts, 0.5)
    eq = oracle_eq("ch2d", mu_f, D_f)
    for b in range(2):
        y, t = y0[b], np.float32(0)
        for d in dts:
            y = O.sifs_step(eq.rhs, y, t, t + np.float32(d), 0.5, eq.fourier_symbol)
            t = t + np.float32(d)
        assert rel_l2(got[b], y) <= 1e-3
        assert np.isfinite(got[b]).all()


def test_ac_steps():
    y0 = make_ic(3, 30, 0.0)
    dts = [5e-6] * 8
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    eq = oracle_eq("ac2d", O.mu_double_well, lambda c: np.ones_like(c))
    plan = SifsPlan("ac2d", N, N, (-N * H / 2,) * 2, (H, H), KAPPA, ("double_well", ()), ("const", (1.0,)))
    sym = torch.from_numpy(fold_symbol(eq.fourier_symbol, 1.0)).cuda()
    out = plan.step(torch.from_numpy(y0).cuda(), dts, sym)
    got = out.cpu().numpy()
    for b in range(3):
        y, t = y0[b], np.float32(0)
        for d in dts:
            y = O.sifs_step(eq.rhs, y, t, t + np.float32(d), 1.0, eq.fourier_symbol)
            t = t + np.float32(d)
        assert rel_l2(got[b], y) <= 1e-5
        assert rel_l2(got[b] - y0[b], y - y0[b]) <= 2e-3


def test_obs_reward_and_control():
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    mu, mob, _, D_f, centre = CASES["log_degenerate"]
    B = 4
    y0 = make_ic(B, 40, centre)
    dom = O.Domain((N, N), ((-N * H / 2, N * H / 2),) * 2)
    X, Y = dom.mesh(np.float32)
    ctrl = np.zeros((B, 8), np.float32)
    ctrl[:, 0] = [0.0, 0.25, -0.5, 0.1]
    ctrl[:, 1] = [0.0, 0.5, 1.0, -0.7]
    ctrl[:, 2] = [0.0, 0.1, -0.2, 0.3]
    ctrl[:, 3] = [0.0, -0.3, 0.2, 0.05]
    ctrl[:, 4] = [0.1, 0.1, 0.05, 0.2]
    dts = [1e-6] * 4
    plan = SifsPlan("ch2d", N, N, (-N * H / 2,) * 2, (H, H), KAPPA, mu, mob)
    eq0 = oracle_eq("ch2d", O.mu_double_well, D_f)
    sym = torch.from_numpy(fold_symbol(eq0.fourier_symbol, 0.5)).cuda()
    obs = torch.empty((B, N, N), dtype=torch.uint8, device="cuda")
    rew = torch.empty((B, 2), dtype=torch.float32, device="cuda")
    out = plan.step(
        torch.from_numpy(y0).cuda(), dts, sym,
        ctrl=torch.from_numpy(ctrl).cuda(), obs=obs, obs_range=(0.0, 1.0), reward=rew,
    )
    got, obs, rew = out.cpu().numpy(), obs.cpu().numpy(), rew.cpu().numpy()
    for b in range(B):
        w = 3.0 + ctrl[b, 0]
        bump = ctrl[b, 1] * np.exp(-((X - ctrl[b, 2]) ** 2 + (Y - ctrl[b, 3]) ** 2) / (2 * ctrl[b, 4] ** 2))
        eq = O.CahnHilliardPeriodic(dom, KAPPA, lambda c: O.mu_log(c, w), D_f, "fd", np.float32, forcing=bump.astype(np.float32))
        y, t = y0[b], np.float32(0)
        for d in dts:
            y = O.sifs_step(eq.rhs, y, t, t + np.float32(d), 0.5, eq.fourier_symbol)
            t = t + np.float32(d)
        assert rel_l2(got[b], y) <= 1e-5
        assert rel_l2(got[b] - y0[b], y - y0[b]) <= 2e-3
        ref_obs = O.quantise_obs(got[b])[0]
        assert np.abs(obs[b].astype(int) - ref_obs.astype(int)).max() <= 1
        assert (obs[b] == ref_obs).mean() > 0.999
        np.testing.assert_allclose(rew[b, 0], got[b].astype(np.float64).mean(), rtol=1e-5)
        np.testing.assert_allclose(rew[b, 1], got[b].astype(np.float64).var(), rtol=1e-4)


@pytest.mark.parametrize("B", [5, 2050])
def test_host_buffer_entry_point_matches_device_path(B):
    """pdeopt_sifs_step_batched_host (pipelined chunks, host buffers) == the device-pointer call,
    bit for bit, including the observation / reward epilogue and a ragged last chunk."""
    import torch

    from pde_opt_b200.fused import SifsPlan, fold_symbol

    N_, H_ = 128, 0.01
    plan = SifsPlan("ch2d", N_, N_, (-N_ * H_ / 2,) * 2, (H_, H_), 0.002, ("log", (3.0,)), ("degenerate", ()))
    k = np.fft.fftfreq(N_, H_)
    k2 = -((2 * np.pi * k[:, None]) ** 2 + (2 * np.pi * k[None, :]) ** 2)
    sym = fold_symbol((0.002 * k2**2).astype(np.complex64), 0.5)
    rng = np.random.default_rng(B)
    y0 = np.clip(0.5 + 0.01 * rng.normal(size=(B, N_, N_)), 0, 1).astype(np.float32)
    ctrl = np.zeros((B, 8), np.float32)
    ctrl[:, 0] = rng.uniform(-0.2, 0.2, B)
    ctrl[:, 1] = rng.uniform(-0.5, 0.5, B)
    ctrl[:, 2:4] = rng.uniform(-0.4, 0.4, (B, 2))
    ctrl[:, 4] = rng.uniform(0.05, 0.2, B)
    dts = [1e-6, 1e-6, 2e-6]
    y1h, obs_h, rew_h = np.empty_like(y0), np.empty((B, N_, N_), np.uint8), np.empty((B, 2), np.float32)
    plan.step_host(y0, dts, sym, ctrl=ctrl, obs=obs_h, reward=rew_h, out=y1h)
    obs_d = torch.empty((B, N_, N_), dtype=torch.uint8, device="cuda")
    rew_d = torch.empty((B, 2), dtype=torch.float32, device="cuda")
    y1d = plan.step(torch.from_numpy(y0).cuda(), dts, torch.from_numpy(sym).cuda(), ctrl=torch.from_numpy(ctrl).cuda(), obs=obs_d, reward=rew_d)
    np.testing.assert_array_equal(y1h, y1d.cpu().numpy())
    np.testing.assert_array_equal(obs_h, obs_d.cpu().numpy())
    np.testing.assert_array_equal(rew_h, rew_d.cpu().numpy())


@pytest.mark.parametrize("kind", ["ch", "ac"])
def test_rhs_fourier_and_steps_match_oracle(kind):
    """derivs='fourier' (cahn_hilliard.py:82-87, allen_cahn.py:74-79): the complex-arithmetic fused
    kernel against the oracle's rhs_fourier, for the RHS itself and for semi-implicit steps."""
    import torch

    from oracle import pde_oracle as O
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AllenCahn2DPeriodic, CahnHilliard2DPeriodic
    from pde_opt_b200.functions import DegenerateMobility, LogRegular, OnePlusSquare
    from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral

    N_, H_ = 128, 0.01
    box = ((-N_ * H_ / 2, N_ * H_ / 2),) * 2
    dom, odom = Domain((N_, N_), box, "dimensionless"), O.Domain((N_, N_), box)
    if kind == "ch":
        eq = CahnHilliard2DPeriodic(dom, 0.002, LogRegular(3.0), DegenerateMobility(), derivs="fourier")
        oeq = O.CahnHilliardPeriodic(odom, 0.002, lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, "fourier", np.float32)
        o64 = O.CahnHilliardPeriodic(odom, 0.002, lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, "fourier", np.float64)
        A, dt = 0.5, 1e-6
    else:
        eq = AllenCahn2DPeriodic(dom, 0.002, LogRegular(3.0), OnePlusSquare(), derivs="fourier")
        oeq = O.AllenCahn2DPeriodic(odom, 0.002, lambda c: O.mu_log(c, 3.0), lambda c: 1 + c**2, "fourier", np.float32)
        o64 = O.AllenCahn2DPeriodic(odom, 0.002, lambda c: O.mu_log(c, 3.0), lambda c: 1 + c**2, "fourier", np.float64)
        A, dt = 1.0, 5e-6
    assert eq.fused
    y0 = np.stack([np.clip(0.5 + 0.05 * np.random.default_rng(s).normal(size=(N_, N_)), 0.05, 0.95) for s in range(3)]).astype(np.float32)
    f = eq.rhs(torch.from_numpy(y0).cuda()).cpu().numpy()
    for b in range(3):
        want = o64.rhs(y0[b].astype(np.float64))
        assert np.linalg.norm(f[b] - want) / np.linalg.norm(want) <= 5e-5
    solver = SemiImplicitFourierSpectral(A, eq.fourier_symbol, eq.fft, eq.ifft)
    times = (np.arange(9, dtype=np.float32) * np.float32(dt)).astype(np.float32)
    got = solver.rollout(ODETerm(eq), times, torch.from_numpy(y0).cuda()).cpu().numpy()
    for b in range(3):
        y = y0[b]
        for a, bb in zip(times[:-1], times[1:]):
            y = O.sifs_step(oeq.rhs, y, a, bb, A, oeq.fourier_symbol)
        assert np.linalg.norm(got[b] - y) / np.linalg.norm(y) <= 1e-5
        assert np.linalg.norm((got[b] - y0[b]) - (y - y0[b])) / np.linalg.norm(y - y0[b]) <= 2e-3
