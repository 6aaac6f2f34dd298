"""PDEEnv / PDEVecEnv — reinforcement-learning environments over the fused stepper.

`PDEEnv` mirrors pde_opt/pde_env.py:22-317 (same 16 constructor arguments, reset/step
contract, observation / action spaces).  `PDEVecEnv` is the batched form the B200 path is
built for: B environments advance in one launch per env-step with the
uint8 observation and the (mean, variance) reward computed in the kernel epilogue."""
from typing import Any, Callable, Dict, Optional

import numpy as np
import torch

from .schedule import constant_step_times
from .solvers import ODETerm
from .utils import check_equation_solver_compatibility, prepare_solver_params

try:  # gymnasium is optional (not installed in the build image)
    import gymnasium as gym
    from gymnasium import spaces

    _EnvBase = gym.Env
except Exception:  # pragma: no cover - exercised when gymnasium is absent
    gym = None

    class _Box:
        def __init__(self, low, high, shape, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype
            self._rng = np.random.default_rng()

        def sample(self):
            return self._rng.uniform(self.low, self.high, self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class _Discrete:
        def __init__(self, n):
            self.n = int(n)
            self._rng = np.random.default_rng()

        def sample(self):
            return int(self._rng.integers(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class spaces:  # noqa: N801 - mirrors gymnasium.spaces
        Box = _Box
        Discrete = _Discrete

    class _EnvBase:
        pass


class PDEEnv(_EnvBase):
    """Single-environment API of the reference (pde_env.py:22-317)."""

    def __init__(self, equation_type, domain, solver_type, end_time: float, step_dt: float, numeric_dt: float,
                 state_to_observation_func: Callable, reward_function: Callable, reset_func: Callable,
                 reset_control_value, update_control_value: Callable, update_control_parameter: Callable,
                 action_space_config: Dict[str, Any], static_equation_parameters: Dict[str, Any],
                 control_equation_parameter_name: str, solver_parameters: Dict[str, Any]):
        super().__init__()
        self.equation_type, self.domain, self.solver_type = equation_type, domain, solver_type
        check_equation_solver_compatibility(self.solver_type, self.equation_type)  # pde_env.py:107
        self.end_time, self.step_dt, self.numeric_dt = end_time, step_dt, numeric_dt
        self.reward_function = reward_function
        self.reset_func = reset_func
        self.state_to_observation_func = state_to_observation_func
        self.observation_space = spaces.Box(low=0.0, high=255.0, shape=(1, *self.domain.points), dtype=np.uint8)
        self._setup_action_space(action_space_config)
        self.reset_control_value = reset_control_value
        self.update_control_value = update_control_value
        self.update_control_parameter = update_control_parameter
        self.static_equation_parameters = static_equation_parameters
        self.control_equation_parameter_name = control_equation_parameter_name
        self.solver_parameters = solver_parameters
        # the step grid of diffeqsolve(t0=0, t1=step_dt, dt0=numeric_dt) is the same every env step
        self._times = constant_step_times(0.0, step_dt, numeric_dt, np.float32, 1_000_000)

    def _setup_action_space(self, config):  # pde_env.py:140-170
        if config.get("type", "continuous") == "discrete":
            self.action_space = spaces.Discrete(config.get("num_actions", 5))
            self._action_to_direction = config.get("action_mapping", {})
        else:
            self.action_space = spaces.Box(
                low=config.get("low", -1.0), high=config.get("high", 1.0), shape=config.get("shape", (2,))
            )
            self._action_to_direction = None

    def _get_obs(self):
        return self.state_to_observation_func(self._state)

    def _get_info(self):
        return {}

    def _terminate(self):
        return self._time >= self.end_time

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):  # pde_env.py:217-242
        self._state = self.reset_func(self.domain, seed=seed) if seed is not None else self.reset_func(self.domain)
        if not torch.is_tensor(self._state):
            self._state = torch.as_tensor(np.asarray(self._state, dtype=np.float32))
        self._state = self._state.to("cuda", torch.float32).contiguous()
        self._time = 0.0
        self._control_value = self.reset_control_value
        return self._get_obs(), self._get_info()

    def step(self, action):  # pde_env.py:244-317
        offset = action if not self._action_to_direction else self._action_to_direction[action]
        old = self._control_value
        self._control_value = self.update_control_value(offset, old)
        control_parameter = self.update_control_parameter(old, self._control_value)
        params = {**self.static_equation_parameters, self.control_equation_parameter_name: control_parameter}
        eq = self.equation_type(domain=self.domain, **params)
        solver = self.solver_type(**prepare_solver_params(self.solver_type, self.solver_parameters, eq))
        self._state = solver.rollout(ODETerm(eq), self._times, self._state)
        self._time += self.step_dt
        obs = self._get_obs()
        reward = self.reward_function(self._state)
        return obs, reward, self._terminate(), False, self._get_info()


class NoiseReset:
    """Batched device-side reset: clip(mean + std * N(0,1), lo, hi) for any number of environments in one
    call (the distribution of the reference notebooks' reset functions, e.g. optimize_nn_script.py:40).
    Called as reset_func(domain, seed=..., batch=n, device=...) -> [n, *points] float32 CUDA tensor."""

    batched = True

    def __init__(self, mean=0.5, std=0.01, lo=0.0, hi=1.0):
        self.mean, self.std, self.lo, self.hi = float(mean), float(std), float(lo), float(hi)

    def __call__(self, domain, seed=None, batch=1, device="cuda"):
        g = torch.Generator(device=device)
        if seed is not None:
            g.manual_seed(int(seed))
        else:
            g.seed()
        y = torch.randn((batch, *domain.points), device=device, generator=g, dtype=torch.float32)
        return y.mul_(self.std).add_(self.mean).clamp_(self.lo, self.hi)


class StateReset:
    """Batched reset to (a perturbation of) a given state, e.g. the GPE ground state (pde_opt/data/ground_state.npy)."""

    batched = True

    def __init__(self, state, rel_noise=0.0):
        self.state, self.rel_noise = np.asarray(state, dtype=np.float32), float(rel_noise)
        self._dev = {}

    def __call__(self, domain, seed=None, batch=1, device="cuda"):
        k = str(device)
        if k not in self._dev:
            self._dev[k] = torch.from_numpy(self.state).to(device)
        y = self._dev[k].unsqueeze(0).repeat(batch, *([1] * self.state.ndim)).contiguous()
        if self.rel_noise:
            g = torch.Generator(device=device)
            g.manual_seed(int(seed)) if seed is not None else g.seed()
            y.mul_(1.0 + self.rel_noise * torch.randn(y.shape, device=device, generator=g))
        return y


class PDEVecEnv:
    """B independent environments stepped by one fused launch per env-step: the batched form of
    PDEEnv.reset / PDEEnv.step (pde_env.py:217-317) with the gymnasium VectorEnv call signatures
    (reset(seed, options) -> (obs, infos); step(actions) -> (obs, rewards, terminations, truncations, infos);
    num_envs, single_observation_space / single_action_space).

    Equations: CahnHilliard2DPeriodic / AllenCahn2DPeriodic with derivs='fd' (uint8 observation and (mean,
    variance) reward from the kernel epilogue), GPE2DTSControl (Strang; observation = quantised density |psi|^2,
    rewards "density_var" | "vortices" = -number of quantised vortices, rl_utils.py:19-84 on the batch |
    ("probe", i, j)), AdvectionDiffusion2D (fused forward kernel).

    Control: `action_to_control(actions, ctrl)` writes the C-ABI control block (include/pdeopt_b200.h: [B, 8] for
    the phase-field and GPE kernels, [B, nseg, 4] = (cx, cy, p0, p1) for advection-diffusion) on the device.
    Everything per environment lives on the device: state, control, time, termination and failure flags.
    Environments whose time reaches `end_time` (pde_env.py:206-215) or whose state has become non-finite
    (pdeopt_plan_set_nonfinite_flags / pdeopt_nonfinite_flags; `info["nonfinite"]`) are reset in one batched call
    when `auto_reset` is on.  `reset_func(domain, seed=..., batch=n, device=...)` with attribute `batched = True`
    (NoiseReset, StateReset) produces n initial states at once; a reference-style reset_func(domain[, seed]) is
    still accepted and looped over."""

    def __init__(self, equation, solver, num_envs, end_time, step_dt, numeric_dt, reset_func,
                 action_to_control: Optional[Callable] = None, obs_range=(0.0, 1.0), reward="var",
                 device="cuda", auto_reset=True, action_space=None, ad_segments=1):
        kind = getattr(equation, "_kind", None)
        name = type(equation).__name__
        if kind in ("ch2d", "ac2d"):
            if not getattr(equation, "fused", False) or getattr(equation, "derivs", "fd") != "fd":
                raise NotImplementedError("PDEVecEnv: phase-field equations need derivs='fd' and enumerated closures")
            self.kind = "phase"
        elif name == "GPE2DTSControl":
            if not getattr(equation, "fused", False):
                raise NotImplementedError("PDEVecEnv: GPE needs an enumerated `lights` (None or GaussianLight); per-env spots go through the control block")
            self.kind = "gpe"
        elif name == "AdvectionDiffusion2D":
            self.kind = "ad"
        else:
            raise NotImplementedError(f"PDEVecEnv: no batched stepper for {name}")
        self.eq, self.solver, self.B = equation, solver, int(num_envs)
        self.num_envs = self.B
        self.end_time, self.step_dt, self.numeric_dt = end_time, step_dt, numeric_dt
        self.reset_func = reset_func
        self.action_to_control = action_to_control
        self.obs_range, self.reward_kind, self.auto_reset = obs_range, reward, auto_reset
        self.device = torch.device(device)
        nx, ny = equation.domain.points
        self.single_observation_space = spaces.Box(low=0.0, high=255.0, shape=(1, nx, ny), dtype=np.uint8)
        self.observation_space = self.single_observation_space
        self.single_action_space = action_space if action_space is not None else spaces.Box(low=-1.0, high=1.0, shape=(8,))
        self.action_space = self.single_action_space
        self._times = constant_step_times(0.0, step_dt, numeric_dt, np.float32, 1_000_000)
        self._terms = ODETerm(equation)
        shape = (self.B, nx, ny, 2) if self.kind == "gpe" else (self.B, nx, ny)
        self.state = torch.empty(shape, dtype=torch.float32, device=self.device)
        self._next = torch.empty_like(self.state)
        self.obs = torch.empty((self.B, 1, nx, ny), dtype=torch.uint8, device=self.device)
        self.stats = torch.empty((self.B, 2), dtype=torch.float32, device=self.device)
        self.flags = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        if self.kind == "ad":
            self.ctrl = torch.zeros((self.B, int(ad_segments), 4), dtype=torch.float32, device=self.device)
        else:
            self.ctrl = torch.zeros((self.B, 8), dtype=torch.float32, device=self.device)
        self._ctrl_default()
        self.time = torch.zeros(self.B, dtype=torch.float64, device=self.device)
        self._episode = 0

    # ---- helpers ----
    def _ctrl_default(self, mask=None):
        sel = slice(None) if mask is None else mask
        self.ctrl[sel] = 0.0
        if self.kind == "ad":
            v = self.eq.velocity
            self.ctrl[sel] = torch.tensor(v.control_row(), dtype=torch.float32, device=self.device)
        else:
            self.ctrl[sel, 4] = 1.0

    def _initial_states(self, n, seed):
        rf = self.reset_func
        if getattr(rf, "batched", False):
            return rf(self.eq.domain, seed=seed, batch=n, device=self.device).to(torch.float32)
        out = []
        for b in range(n):
            s = rf(self.eq.domain, seed=(None if seed is None else seed + b)) if seed is not None else rf(self.eq.domain)
            out.append(torch.as_tensor(np.asarray(s, dtype=np.float32)) if not torch.is_tensor(s) else s.to(torch.float32))
        return torch.stack(out).to(self.device)

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        self.state.copy_(self._initial_states(self.B, seed))
        self.time.zero_()
        self.flags.zero_()
        self._ctrl_default()
        return self._observe(), {}

    def _field(self):
        if self.kind == "gpe":
            return self.state[..., 0] ** 2 + self.state[..., 1] ** 2
        return self.state

    def _observe(self):
        lo, hi = self.obs_range
        q = torch.clamp((self._field() - lo) / (hi - lo), 0.0, 1.0) * 255.0
        self.obs[:, 0] = torch.round(q).to(torch.uint8)
        return self.obs

    def _rollout(self, lo, hi):
        y, out = self.state[lo:hi], self._next[lo:hi]
        if self.kind == "phase":
            self.solver.rollout(self._terms, self._times, y, ctrl=self.ctrl[lo:hi], obs=self.obs[lo:hi], obs_range=self.obs_range,
                                reward=self.stats[lo:hi], out=out, nonfinite=self.flags[lo:hi])
        elif self.kind == "gpe":
            self.solver.rollout(self._terms, self._times, y, ctrl=self.ctrl[lo:hi], out=out)
        else:
            from .adjoint import ad_rollout

            with torch.no_grad():
                out.copy_(ad_rollout(self.eq, y, self.ctrl[lo:hi], self._times, A=float(getattr(self.solver, "A", 1.0))))

    def _epilogue(self):
        """Observation / reward / failure flags of the equations without a fused epilogue (device ops)."""
        if self.kind == "phase":
            return
        import ctypes

        from . import _lib

        self._observe()
        f = self._field()
        self.stats[:, 0] = f.mean(dim=(1, 2))
        self.stats[:, 1] = f.var(dim=(1, 2), unbiased=False)
        n_per = int(self.state[0].numel())
        with _lib.device_of(self.state):
            _lib.check(_lib.load().pdeopt_nonfinite_flags(ctypes.c_void_p(self.state.data_ptr()), self.B, n_per,
                                                          ctypes.c_void_p(self.flags.data_ptr()), _lib.stream_ptr(self.state)))

    def _reward(self):
        rk = self.reward_kind
        if isinstance(rk, tuple) and rk[0] == "probe":
            return self._field()[:, rk[1], rk[2]]  # point probe of the new state
        if rk == "vortices":
            from .rl_utils import vortex_counts

            return -vortex_counts(self.state)[:, 0].to(torch.float32)
        if rk in ("var", "density_var"):
            return self.stats[:, 1]
        return self.stats[:, 0]

    def step(self, actions, obs_host=None, stats_host=None, chunks=4):
        """One env step of all B environments.  With pinned host tensors `obs_host` [B, 1, nx, ny] uint8
        and `stats_host` [B, 2] float32 the batch is stepped in `chunks` slices on separate streams, so
        that the device->host copy of one slice's observation overlaps the next slice's kernel; the
        call returns when everything has landed on the host."""
        if self.action_to_control is not None:
            self.action_to_control(actions, self.ctrl)
        if obs_host is None or self.kind != "phase":
            self._rollout(0, self.B)
            self.state, self._next = self._next, self.state
            self._epilogue()
            if obs_host is not None:
                obs_host.copy_(self.obs, non_blocking=True)
                if stats_host is not None:
                    stats_host.copy_(self.stats, non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
        else:
            chunks = max(1, min(int(chunks), self.B))
            if not hasattr(self, "_streams") or len(self._streams) != chunks:
                self._streams = [torch.cuda.Stream(device=self.device) for _ in range(chunks)]
            main = torch.cuda.current_stream(self.device)
            step = (self.B + chunks - 1) // chunks
            for c, st in enumerate(self._streams):
                lo, hi = c * step, min(self.B, (c + 1) * step)
                if lo >= hi:
                    break
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    self._rollout(lo, hi)
                    obs_host[lo:hi].copy_(self.obs[lo:hi], non_blocking=True)
                    if stats_host is not None:
                        stats_host[lo:hi].copy_(self.stats[lo:hi], non_blocking=True)
            for st in self._streams:
                main.wait_stream(st)
            main.synchronize()
            self.state, self._next = self._next, self.state
        self.time += self.step_dt
        reward = self._reward()
        failed = self.flags != 0
        terminated = (self.time >= self.end_time) | failed
        truncated = torch.zeros_like(terminated)
        info = {"nonfinite": failed}
        if self.auto_reset:
            self._auto_reset(terminated)
        return self.obs, reward, terminated, truncated, info

    def _auto_reset(self, mask):
        """Resets the masked environments in one batched call; no host loop over B (one scalar sync for the count)."""
        n = int(mask.sum().item())
        if n == 0:
            return
        self._episode += 1
        self.state[mask] = self._initial_states(n, None if self._episode is None else 7919 * self._episode)
        self.time[mask] = 0.0
        self.flags[mask] = 0
        self._ctrl_default(mask)
