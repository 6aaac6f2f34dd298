"""GPU parity of the reference-facing API (equations / solver protocol / PDEModel / PDEEnv)
against the NumPy oracle."""
import numpy as np
import pytest

from oracle import pde_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N, H, KAPPA = 128, 0.01, 0.002


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def ic(seed, centre=0.5):
    return np.clip(centre + 0.01 * np.random.default_rng(seed).normal(size=(N, N)), 0, 1).astype(np.float32)


def odom():
    return O.Domain((N, N), ((-N * H / 2, N * H / 2),) * 2)


def pdom():
    from pde_opt_b200 import Domain

    return Domain((N, N), ((-N * H / 2, N * H / 2),) * 2, "dimensionless")


def test_rhs_matches_oracle_tightly():
    """eq.rhs (pdeopt_rhs_batched) vs cahn_hilliard.py:89-109 / allen_cahn.py:81-84 restated."""
    from pde_opt_b200.equations import AllenCahn2DPeriodic, CahnHilliard2DPeriodic

    u = np.stack([ic(s) for s in range(3)])
    cases = [
        (CahnHilliard2DPeriodic, lambda c: c**3 - c, lambda c: 0 * c + 1.0, "ch", 3e-6),
        (CahnHilliard2DPeriodic, lambda c: c**3 - c, lambda c: 1 + c**2, "ch", 3e-6),
        (CahnHilliard2DPeriodic, lambda c: np.log(c / (1.0 - c)) + 3.0 * (1.0 - 2.0 * c), lambda c: (1 - c) * c, "ch", 2e-4),
        (AllenCahn2DPeriodic, lambda c: c**3 - c, lambda c: 0 * c + 1.0, "ac", 3e-6),
    ]
    for cls, mu, mob, kind, tol in cases:
        eq = cls(pdom(), KAPPA, mu, mob)
        got = eq.rhs(torch.from_numpy(u).cuda(), 0.0).cpu().numpy()
        if kind == "ch":
            oeq = O.CahnHilliardPeriodic(odom(), KAPPA, mu, mob, "fd", np.float32)
        else:
            oeq = O.AllenCahn2DPeriodic(odom(), KAPPA, mu, mob, "fd", np.float32)
        for b in range(3):
            # float64 oracle is the truth; float32 oracle gives the noise floor of the formula itself
            ref64 = (O.CahnHilliardPeriodic if kind == "ch" else O.AllenCahn2DPeriodic)(odom(), KAPPA, mu, mob, "fd", np.float64).rhs(u[b].astype(np.float64))
            floor = rel_l2(oeq.rhs(u[b]), ref64)
            assert rel_l2(got[b], ref64) <= max(tol, 4 * floor), (kind, b)
    single = eq.rhs(torch.from_numpy(u[0]).cuda(), 0.0)
    assert single.shape == (N, N)


def test_solver_step_fused_and_unfused_paths():
    """SemiImplicitFourierSpectral.step return tuple (solvers.py:56-70) on both paths."""
    from pde_opt_b200.equations import CahnHilliard2DPeriodic
    from pde_opt_b200.solvers import RESULTS, ODETerm, SemiImplicitFourierSpectral

    eq = CahnHilliard2DPeriodic(pdom(), KAPPA, lambda c: c**3 - c, lambda c: 0 * c + 1.0)
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft, with_error=True)
    y0 = torch.from_numpy(ic(3, 0.0)).cuda()
    oeq = O.CahnHilliardPeriodic(odom(), KAPPA, O.mu_double_well, lambda c: np.ones_like(c), "fd", np.float32)
    ref, ref_err = O.sifs_step(oeq.rhs, ic(3, 0.0), np.float32(0.0), np.float32(2e-6), 0.5, oeq.fourier_symbol, with_error=True)
    for terms in (ODETerm(eq), ODETerm(lambda t, y, args: eq.rhs(y, t))):
        y1, y_err, dense, state, result = solver.step(terms, 0.0, 2e-6, y0, None, None, False)
        assert state is None and result == RESULTS.successful
        assert dense["y0"] is y0 and dense["y1"] is y1
        assert rel_l2(y1.cpu().numpy(), ref) <= 1e-5
        assert rel_l2(y1.cpu().numpy() - ic(3, 0.0), ref - ic(3, 0.0)) <= 2e-3
        assert rel_l2(y_err.cpu().numpy(), ref_err) <= 2e-3


def test_pde_model_solve_saveat_interpolation():
    """PDEModel.solve (pde_model.py:68-136): constant steps + SaveAt(ts) linear interpolation."""
    from pde_opt_b200.equations import CahnHilliard2DPeriodic
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    model = PDEModel(CahnHilliard2DPeriodic, pdom(), SemiImplicitFourierSpectral)
    mu = lambda c: np.log(c / (1.0 - c)) + 3.0 * (1.0 - 2.0 * c)  # noqa: E731
    params = {"kappa": KAPPA, "mu": mu, "D": lambda c: np.ones_like(c)}
    y0 = ic(7)
    ts = np.linspace(0.0, 2.0e-4, 7)
    ys = model.solve(params, torch.from_numpy(y0).cuda(), ts, {"A": 0.5}, dt0=1e-6, max_steps=1_000_000)
    assert tuple(ys.shape) == (7, N, N)
    oeq = O.CahnHilliardPeriodic(odom(), KAPPA, lambda c: O.mu_log(c, 3.0), lambda c: np.ones_like(c), "fd", np.float32)
    ref = O.integrate(lambda y, a, b: O.sifs_step(oeq.rhs, y, a, b, 0.5, oeq.fourier_symbol), y0, 0.0, 2.0e-4, 1e-6,
                      save_ts=ts.astype(np.float32))
    got = ys.cpu().numpy()
    np.testing.assert_array_equal(got[0], y0)
    for i in range(1, 7):
        assert rel_l2(got[i], ref[i]) <= 1e-4, i
    # batched solve == per-trajectory solve (the reference vmaps solve over trajectories, pde_model.py:266-268)
    yb = np.stack([ic(8), ic(9)])
    ysb = model.solve(params, torch.from_numpy(yb).cuda(), ts[:3], {"A": 0.5}, dt0=1e-6).cpu().numpy()
    ys8 = model.solve(params, torch.from_numpy(yb[0]).cuda(), ts[:3], {"A": 0.5}, dt0=1e-6).cpu().numpy()
    assert tuple(ysb.shape) == (3, 2, N, N)
    # two environments share one complex FFT, so a trajectory depends on its batch partner at the ulp level
    np.testing.assert_allclose(ysb[:, 0], ys8, rtol=0, atol=2e-6)


def test_pde_model_incompatible_pair_raises():
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    class Eq:
        pass

    with pytest.raises(ValueError):
        PDEModel(Eq, pdom(), SemiImplicitFourierSpectral)


def test_pde_env_reset_step_contract():
    """PDEEnv.reset/step (pde_env.py:217-317) with the notebook's step_dt / numeric_dt ratio."""
    from pde_opt_b200.equations import CahnHilliard2DPeriodic
    from pde_opt_b200.pde_env import PDEEnv
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    def reset_func(domain, seed=0):
        return ic(seed)

    def to_obs(state):
        return O.quantise_obs(state.cpu().numpy())

    env = PDEEnv(
        equation_type=CahnHilliard2DPeriodic, domain=pdom(), solver_type=SemiImplicitFourierSpectral,
        end_time=1.0e-4, step_dt=5.0e-5, numeric_dt=1e-6,
        state_to_observation_func=to_obs, reward_function=lambda s: float(s.var()),
        reset_func=reset_func, reset_control_value=3.0,
        update_control_value=lambda off, old: old + off,
        update_control_parameter=lambda old, new: (lambda c, w=new: np.log(c / (1.0 - c)) + w * (1.0 - 2.0 * c)),
        action_space_config={"type": "discrete", "num_actions": 3, "action_mapping": {0: -0.1, 1: 0.0, 2: 0.1}},
        static_equation_parameters={"kappa": KAPPA, "D": lambda c: (1 - c) * c},
        control_equation_parameter_name="mu", solver_parameters={"A": 0.5},
    )
    obs, info = env.reset(seed=11)
    assert obs.shape == (1, N, N) and obs.dtype == np.uint8 and info == {}
    assert env.observation_space.shape == (1, N, N)
    y = ic(11)
    w = 3.0
    for action, done_expected in ((2, False), (0, True)):
        obs, reward, terminated, truncated, info = env.step(action)
        w = w + {0: -0.1, 1: 0.0, 2: 0.1}[action]
        oeq = O.CahnHilliardPeriodic(odom(), KAPPA, lambda c, w=w: O.mu_log(c, w), lambda c: (1 - c) * c, "fd", np.float32)
        y = O.integrate(lambda yy, a, b: O.sifs_step(oeq.rhs, yy, a, b, 0.5, oeq.fourier_symbol), y, 0.0, 5.0e-5, 1e-6)[0]
        assert rel_l2(env._state.cpu().numpy(), y) <= 1e-4
        assert terminated == done_expected and truncated is False and info == {}
        np.testing.assert_allclose(reward, y.astype(np.float64).var(), rtol=1e-3)
        assert obs.shape == (1, N, N)


def test_vec_env_matches_single_envs():
    from pde_opt_b200.equations import CahnHilliard2DPeriodic
    from pde_opt_b200.functions import DegenerateMobility, LogRegular
    from pde_opt_b200.pde_env import PDEVecEnv
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    eq = CahnHilliard2DPeriodic(pdom(), KAPPA, LogRegular(3.0), DegenerateMobility())
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    B = 5

    def act(actions, ctrl):
        ctrl[:, 0] = torch.as_tensor(actions, dtype=torch.float32, device=ctrl.device)

    env = PDEVecEnv(eq, solver, B, end_time=1.0, step_dt=1.6e-5, numeric_dt=1e-6, reset_func=lambda d, seed=None: ic(seed or 0),
                    action_to_control=act)
    obs, _ = env.reset(seed=100)
    assert tuple(obs.shape) == (B, 1, N, N) and obs.dtype == torch.uint8
    actions = np.array([0.0, 0.1, -0.1, 0.2, 0.05], np.float32)
    obs, reward, term, trunc, _ = env.step(actions)
    times = O.constant_step_schedule(0.0, 1.6e-5, 1e-6, np.float32)
    for b in range(B):
        oeq = O.CahnHilliardPeriodic(odom(), KAPPA, lambda c, w=3.0 + actions[b]: O.mu_log(c, w), lambda c: (1 - c) * c, "fd", np.float32)
        y = ic(100 + b)
        for a, bb in zip(times[:-1], times[1:]):
            y = O.sifs_step(oeq.rhs, y, a, bb, 0.5, oeq.fourier_symbol)
        assert rel_l2(env.state[b].cpu().numpy(), y) <= 1e-5
        np.testing.assert_allclose(float(reward[b]), y.astype(np.float64).var(), rtol=1e-3)
        assert (obs[b, 0].cpu().numpy() == O.quantise_obs(env.state[b].cpu().numpy())[0]).mean() > 0.999
    assert not term.any() and not trunc.any()
    # host delivery: slices on separate streams, observation / reward copies overlapped with the kernels;
    # identical results, on the host when the call returns
    env2 = PDEVecEnv(eq, solver, B, end_time=1.0, step_dt=1.6e-5, numeric_dt=1e-6, reset_func=lambda d, seed=None: ic(seed or 0),
                     action_to_control=act)
    env2.reset(seed=100)
    obs_h = torch.empty((B, 1, N, N), dtype=torch.uint8).pin_memory()
    st_h = torch.empty((B, 2), dtype=torch.float32).pin_memory()
    env2.step(actions, obs_host=obs_h, stats_host=st_h, chunks=2)
    assert torch.equal(env2.state, env.state)
    assert torch.equal(obs_h, obs.cpu()) and torch.equal(st_h, env.stats.cpu())


def test_pde_env_with_advection_diffusion_and_gpe():
    """PDEEnv is generic over (equation_type, solver_type) as in the reference (pde_env.py:43-61):
    the recovered advection-diffusion equation with a moving velocity centre as the control
    (the deleted AdvectionDiffusionEnv of notebooks/test_pde_RL.ipynb:129), and the GPE with
    StrangSplitting."""
    import torch

    from oracle import pde_oracle as O
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AdvectionDiffusion2D, GPE2DTSControl
    from pde_opt_b200.functions import GaussianLight, GaussianVelocity
    from pde_opt_b200.pde_env import PDEEnv
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral, StrangSplitting

    n, h = 128, 0.02
    box = ((-n * h / 2, n * h / 2),) * 2
    dom, odom = Domain((n, n), box, "dimensionless"), O.Domain((n, n), box)
    u0 = (0.5 + 0.01 * np.random.default_rng(0).normal(size=(n, n))).astype(np.float32)
    env = PDEEnv(
        AdvectionDiffusion2D, dom, SemiImplicitFourierSpectral, end_time=0.01, step_dt=0.002, numeric_dt=1e-4,
        state_to_observation_func=lambda s: (s.clamp(0, 1) * 255).round().to(torch.uint8)[None],
        reward_function=lambda s: float(s.var()), reset_func=lambda d, seed=None: u0,
        reset_control_value=np.array([0.0, 0.0]), update_control_value=lambda off, old: old + np.asarray(off),
        update_control_parameter=lambda old, new: GaussianVelocity(0.1, 0.01, (float(new[0]), float(new[1]))),
        action_space_config={"type": "continuous", "low": -0.1, "high": 0.1, "shape": (2,)},
        static_equation_parameters={"D": 0.1}, control_equation_parameter_name="velocity", solver_parameters={"A": 1.0},
    )
    obs, info = env.reset(seed=1)
    assert obs.shape == (1, n, n) and obs.dtype == torch.uint8
    centre, y = np.zeros(2), u0
    for a in ([0.05, -0.02], [0.03, 0.04]):
        obs, rew, term, trunc, info = env.step(np.asarray(a))
        centre = centre + np.asarray(a)
        oeq = O.AdvectionDiffusion2D(odom, O.gaussian_velocity((0.1, 0.01), tuple(centre)), 0.1, np.float32)
        for ta, tb in zip(env._times[:-1], env._times[1:]):
            y = O.sifs_step(oeq.rhs, y, ta, tb, 1.0, oeq.fourier_symbol)
    got = env._state.cpu().numpy()
    assert np.linalg.norm(got - y) / np.linalg.norm(y) <= 1e-5
    assert not term and trunc is False

    # GPE: the light amplitude is the control
    L_ = 20.0
    gdom = Domain((n, n), ((-L_ / 2, L_ / 2),) * 2, "dimensionless")
    psi = np.exp(-(np.add.outer(np.linspace(-3, 3, n) ** 2, np.linspace(-3, 3, n) ** 2))).astype(np.float32)
    psi0 = np.stack([psi, 0 * psi], -1)
    psi0 = (psi0 / np.sqrt((psi0**2).sum() * (L_ / n) ** 2)).astype(np.float32)
    genv = PDEEnv(
        GPE2DTSControl, gdom, StrangSplitting, end_time=1.0, step_dt=1e-3, numeric_dt=1e-4,
        state_to_observation_func=lambda s: ((s**2).sum(-1).clamp(0, 1) * 255).round().to(torch.uint8)[None],
        reward_function=lambda s: float((s**2).sum(-1).max()), reset_func=lambda d, seed=None: psi0,
        reset_control_value=0.0, update_control_value=lambda off, old: old + off,
        update_control_parameter=lambda old, new: GaussianLight(float(new), 1.0, -1.0, 2.0),
        action_space_config={"type": "discrete", "num_actions": 3, "action_mapping": {0: -1.0, 1: 0.0, 2: 1.0}},
        static_equation_parameters={"k": 100.0, "e": 0.0, "trap_factor": 1.0}, control_equation_parameter_name="lights",
        solver_parameters={"time_scale": -1j},
    )
    genv.reset()
    obs, rew, term, trunc, _ = genv.step(2)
    ogdom = O.Domain((n, n), ((-L_ / 2, L_ / 2),) * 2)
    light = GaussianLight(1.0, 1.0, -1.0, 2.0)
    ogeq = O.GPE2DTSControl(ogdom, 100.0, 0.0, lambda t, x, y: light(t, x, y), 1.0, np.float32)
    yy = psi0
    for ta, tb in zip(genv._times[:-1], genv._times[1:]):
        yy = O.strang_step(ogeq.B_terms, yy, ta, tb, ogeq.A_term, ogeq.dx, -1j)
    assert np.linalg.norm(genv._state.cpu().numpy() - yy) / np.linalg.norm(yy) <= 2e-5


def test_pde_model_solve_with_pid_controller_matches_oracle():
    """PDEModel.solve(..., stepsize_controller=PIDController(rtol, atol)) (pde_model.py:120-134 with an
    adaptive controller; y_error from solvers.py:61-65): same accepted / rejected step counts as the
    oracle's restatement of the controller, states at the save times within the north-star tolerance,
    and the adaptive solution agrees with the constant-step one to the requested accuracy."""
    from pde_opt_b200.equations import AllenCahn2DPeriodic
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.functions import ConstantMobility, DoubleWell
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral
    from pde_opt_b200.stepsize import ConstantStepSize, PIDController

    model = PDEModel(AllenCahn2DPeriodic, pdom(), SemiImplicitFourierSpectral)
    params = dict(kappa=KAPPA, mu=DoubleWell(), R=ConstantMobility(1.0))
    y0 = (0.1 * np.random.default_rng(9).normal(size=(N, N))).astype(np.float32)
    ts = np.linspace(0.0, 2e-3, 5).astype(np.float32)
    rtol, atol = 1e-3, 1e-6
    got = model.solve(params, torch.from_numpy(y0).cuda(), ts, {"A": 1.0}, dt0=1e-6,
                      stepsize_controller=PIDController(rtol, atol)).cpu().numpy()
    stats = dict(model.last_stats)
    oeq = O.AllenCahn2DPeriodic(odom(), KAPPA, O.mu_double_well, lambda c: np.ones_like(c), "fd", np.float32)
    want, acc, rej = O.integrate_adaptive(
        lambda y, a, b: O.sifs_step(oeq.rhs, y, a, b, 1.0, oeq.fourier_symbol, with_error=True), y0, ts[0], ts[-1], 1e-6,
        rtol, atol, save_ts=ts)
    assert np.isfinite(got).all() and got.shape == (5, N, N)
    assert (stats["accepted"], stats["rejected"]) == (acc, rej)
    assert acc < 400  # constant stepping would take 2000 steps of 1e-6
    for i in range(len(ts)):
        assert rel_l2(got[i], want[i]) <= 1e-4
    fine = model.solve(params, torch.from_numpy(y0).cuda(), ts, {"A": 1.0}, dt0=1e-6,
                       stepsize_controller=ConstantStepSize()).cpu().numpy()
    assert rel_l2(got[-1], fine[-1]) <= 0.05
