"""PDEVecEnv over the other equations (GPE with Strang splitting, advection-diffusion), device-side batched
reset, and automatic reset of terminated / non-finite environments (reference: pde_env.py:217-317 per env)."""
import numpy as np
import pytest

from oracle import pde_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))


def test_gpe_vec_env_matches_oracle_per_environment():
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import GPE2DTSControl
    from pde_opt_b200.functions import GaussianLight
    from pde_opt_b200.pde_env import PDEVecEnv, StateReset
    from pde_opt_b200.solvers import StrangSplitting

    n, L_, k = 128, 20.0, 100.0
    box = ((-L_ / 2, L_ / 2),) * 2
    dom, odom = Domain((n, n), box, "dimensionless"), O.Domain((n, n), box)
    X, Y = odom.mesh()
    psi = np.exp(-(X**2 + Y**2) / 8.0)
    psi0 = np.stack([psi, 0 * psi], -1)
    psi0 = (psi0 / np.sqrt((psi0**2).sum() * odom.dx[0] ** 2)).astype(np.float32)
    eq = GPE2DTSControl(dom, k, 0.0, None, 1.0)
    solver = StrangSplitting(eq.A_term, eq.dx, eq.fft, eq.ifft, -1j)
    B = 3
    amps = np.array([0.0, 2.0, -1.5], np.float32)

    def act(actions, ctrl):  # per-environment light spot: amplitude from the action, fixed place and width
        ctrl[:, 1] = torch.as_tensor(actions, dtype=torch.float32, device=ctrl.device)
        ctrl[:, 2], ctrl[:, 3], ctrl[:, 4] = 1.0, -1.0, 2.0

    env = PDEVecEnv(eq, solver, B, end_time=1.0, step_dt=8e-4, numeric_dt=1e-4, reset_func=StateReset(psi0), action_to_control=act,
                    obs_range=(0.0, float((psi0**2).sum(-1).max())), reward="density_var")
    obs, info = env.reset(seed=0)
    assert tuple(obs.shape) == (B, 1, n, n) and env.num_envs == B
    obs, rew, term, trunc, info = env.step(amps)
    times = O.constant_step_schedule(0.0, 8e-4, 1e-4, np.float32)
    for b in range(B):
        light = GaussianLight(float(amps[b]), 1.0, -1.0, 2.0)
        oeq = O.GPE2DTSControl(odom, k, 0.0, lambda t, x, y, light=light: light(t, x, y), 1.0, np.float32)
        y = psi0
        for ta, tb in zip(times[:-1], times[1:]):
            y = O.strang_step(oeq.B_terms, y, ta, tb, oeq.A_term, oeq.dx, -1j)
        got = env.state[b].cpu().numpy()
        assert rel_l2(got, y) <= 2e-5
        dens = (y.astype(np.float64) ** 2).sum(-1)
        np.testing.assert_allclose(float(rew[b]), dens.var(), rtol=1e-3)
    assert not term.any() and not info["nonfinite"].any()
    env.reward_kind = "vortices"
    _, rew, _, _, _ = env.step(amps)
    assert tuple(rew.shape) == (B,) and float(rew.max()) <= 0.0


def test_advection_diffusion_vec_env_matches_oracle():
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AdvectionDiffusion2D
    from pde_opt_b200.functions import GaussianVelocity
    from pde_opt_b200.pde_env import NoiseReset, PDEVecEnv
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    n, h = 128, 0.02
    box = ((-n * h / 2, n * h / 2),) * 2
    dom, odom = Domain((n, n), box, "dimensionless"), O.Domain((n, n), box)
    eq = AdvectionDiffusion2D(dom, GaussianVelocity(0.1, 0.01), 0.1)
    solver = SemiImplicitFourierSpectral(1.0, eq.fourier_symbol, eq.fft, eq.ifft)
    B = 4
    centres = np.array([[0.0, 0.0], [0.05, -0.02], [-0.1, 0.1], [0.2, 0.0]], np.float32)

    def act(actions, ctrl):  # the velocity centre is the control (the deleted AdvectionDiffusionEnv, test_pde_RL.ipynb:129)
        ctrl[:, 0, 0:2] = torch.as_tensor(actions, dtype=torch.float32, device=ctrl.device)

    env = PDEVecEnv(eq, solver, B, end_time=1.0, step_dt=2e-3, numeric_dt=1e-4, reset_func=NoiseReset(0.5, 0.01, 0.0, 1.0), action_to_control=act)
    env.reset(seed=3)
    y0 = env.state.cpu().numpy().copy()
    assert np.isfinite(y0).all() and abs(float(y0.mean()) - 0.5) < 1e-3 and not np.array_equal(y0[0], y0[1])
    obs, rew, term, trunc, info = env.step(centres)
    for b in range(B):
        oeq = O.AdvectionDiffusion2D(odom, O.gaussian_velocity((0.1, 0.01), tuple(float(v) for v in centres[b])), 0.1, np.float32)
        y = y0[b]
        for ta, tb in zip(env._times[:-1], env._times[1:]):
            y = O.sifs_step(oeq.rhs, y, ta, tb, 1.0, oeq.fourier_symbol)
        assert rel_l2(env.state[b].cpu().numpy(), y) <= 1e-5
        np.testing.assert_allclose(float(rew[b]), y.astype(np.float64).var(), rtol=2e-3)


def test_auto_reset_of_terminated_and_non_finite_environments():
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard2DPeriodic
    from pde_opt_b200.functions import DegenerateMobility, LogRegular
    from pde_opt_b200.pde_env import NoiseReset, PDEVecEnv
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    n, h = 128, 0.01
    dom = Domain((n, n), ((-n * h / 2, n * h / 2),) * 2, "dimensionless")
    eq = CahnHilliard2DPeriodic(dom, 0.002, LogRegular(3.0), DegenerateMobility())
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    B = 6
    env = PDEVecEnv(eq, solver, B, end_time=3.2e-5, step_dt=1.6e-5, numeric_dt=1e-6, reset_func=NoiseReset(0.5, 0.01, 0.0, 1.0))
    env.reset(seed=11)
    env.state[2, 7, 9] = float("nan")  # this environment fails in the first step
    obs, rew, term, trunc, info = env.step(None)
    assert info["nonfinite"].cpu().tolist() == [False, False, True, False, False, False]
    assert term.cpu().tolist() == [False, False, True, False, False, False]
    assert torch.isfinite(env.state).all(), "the failed environment must have been reset"
    assert env.time.cpu().tolist() == pytest.approx([1.6e-5, 1.6e-5, 0.0, 1.6e-5, 1.6e-5, 1.6e-5])
    obs, rew, term, trunc, info = env.step(None)  # everybody but the restarted one reaches end_time
    assert term.cpu().tolist() == [True, True, False, True, True, True]
    assert env.time.cpu().tolist() == pytest.approx([0.0, 0.0, 1.6e-5, 0.0, 0.0, 0.0])
    assert torch.isfinite(env.state).all() and not info["nonfinite"].any()
