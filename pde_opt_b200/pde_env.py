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


class PDEVecEnv:
    """B independent environments stepped by one fused launch per env-step.

    The control is the C-ABI control block (include/pdeopt_b200.h): `action_to_control(actions,
    ctrl)` is a user callback that writes the [B, 8] float32 block (device tensor) from the
    actions; observation (uint8, Box(0,255,(1,*points))) and reward (variance by default, the
    reference notebooks' `np.var`; "mean"; or ("probe", i, j) for the value at one grid point) come from
    the kernel epilogue.  Environments whose time
    reaches `end_time` are reset automatically (pde_env.py:206-215 gives the criterion)."""

    def __init__(self, equation, solver, num_envs, end_time, step_dt, numeric_dt, reset_func,
                 action_to_control: Optional[Callable] = None, obs_range=(0.0, 1.0), reward="var",
                 device="cuda", auto_reset=True):
        if getattr(equation, "_kind", None) not in ("ch2d", "ac2d") or not getattr(equation, "fused", False) \
                or getattr(equation, "derivs", "fd") != "fd":
            # the fused observation / reward epilogue exists in the finite-difference phase-field kernels
            # only; failing here is better than returning stale observation buffers
            raise NotImplementedError(
                "PDEVecEnv needs a Cahn-Hilliard / Allen-Cahn 2-D equation with derivs='fd' and enumerated closures; "
                "use PDEEnv (one environment) for the other equation types"
            )
        self.eq, self.solver, self.B = equation, solver, int(num_envs)
        self.end_time, self.step_dt, self.numeric_dt = end_time, step_dt, numeric_dt
        self.reset_func = reset_func
        self.action_to_control = action_to_control
        self.obs_range, self.reward_kind, self.auto_reset = obs_range, reward, auto_reset
        self.device = torch.device(device)
        nx, ny = equation.domain.points
        self.observation_space = spaces.Box(low=0.0, high=255.0, shape=(1, nx, ny), dtype=np.uint8)
        self._times = constant_step_times(0.0, step_dt, numeric_dt, np.float32, 1_000_000)
        self._terms = ODETerm(equation)
        self.state = torch.empty((self.B, nx, ny), dtype=torch.float32, device=self.device)
        self._next = torch.empty_like(self.state)
        self.obs = torch.empty((self.B, 1, nx, ny), dtype=torch.uint8, device=self.device)
        self.stats = torch.empty((self.B, 2), dtype=torch.float32, device=self.device)
        self.ctrl = torch.zeros((self.B, 8), dtype=torch.float32, device=self.device)
        self.ctrl[:, 4] = 1.0
        self.time = np.zeros(self.B, dtype=np.float64)

    def reset(self, seed: Optional[int] = None):
        for b in range(self.B):
            s = self.reset_func(self.eq.domain, seed=(None if seed is None else seed + b))
            self.state[b] = torch.as_tensor(np.asarray(s, dtype=np.float32)).to(self.device)
        self.time[:] = 0.0
        self.ctrl.zero_()
        self.ctrl[:, 4] = 1.0
        return self._observe(), {}

    def _observe(self):
        lo, hi = self.obs_range
        q = torch.clamp((self.state - lo) / (hi - lo), 0.0, 1.0) * 255.0
        self.obs[:, 0] = torch.round(q).to(torch.uint8)
        return self.obs

    def step(self, actions, obs_host=None, stats_host=None, chunks=4):
        """One env step of all B environments.  With pinned host tensors `obs_host` [B, 1, nx, ny] uint8
        and `stats_host` [B, 2] float32 the batch is stepped in `chunks` slices on separate streams, so
        that the device->host copy of one slice's observation overlaps the next slice's kernel; the
        call returns when everything has landed on the host."""
        if self.action_to_control is not None:
            self.action_to_control(actions, self.ctrl)
        if obs_host is None:
            self.solver.rollout(
                self._terms, self._times, self.state, ctrl=self.ctrl, obs=self.obs, obs_range=self.obs_range,
                reward=self.stats, out=self._next,
            )
        else:
            chunks = max(1, min(int(chunks), self.B // 2))
            if not hasattr(self, "_streams") or len(self._streams) != chunks:
                self._streams = [torch.cuda.Stream(device=self.device) for _ in range(chunks)]
            main = torch.cuda.current_stream(self.device)
            step = 2 * ((self.B + 2 * chunks - 1) // (2 * chunks))  # even slices: two environments share a CTA
            for c, st in enumerate(self._streams):
                lo, hi = c * step, min(self.B, (c + 1) * step)
                if lo >= hi:
                    break
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    self.solver.rollout(
                        self._terms, self._times, self.state[lo:hi], ctrl=self.ctrl[lo:hi], obs=self.obs[lo:hi],
                        obs_range=self.obs_range, reward=self.stats[lo:hi], out=self._next[lo:hi],
                    )
                    obs_host[lo:hi].copy_(self.obs[lo:hi], non_blocking=True)
                    if stats_host is not None:
                        stats_host[lo:hi].copy_(self.stats[lo:hi], non_blocking=True)
            for st in self._streams:
                main.wait_stream(st)
            main.synchronize()
        self.state, self._next = self._next, self.state
        self.time += self.step_dt
        if isinstance(self.reward_kind, tuple) and self.reward_kind[0] == "probe":
            reward = self.state[:, self.reward_kind[1], self.reward_kind[2]]  # point probe of the new state
        else:
            reward = self.stats[:, 1] if self.reward_kind == "var" else self.stats[:, 0]
        terminated = self.time >= self.end_time
        if self.auto_reset and terminated.any():
            for b in np.nonzero(terminated)[0]:
                s = self.reset_func(self.eq.domain, seed=None)
                self.state[b] = torch.as_tensor(np.asarray(s, dtype=np.float32)).to(self.device)
                self.time[b] = 0.0
        return self.obs, reward, terminated, np.zeros(self.B, dtype=bool), {}
