"""PDEModel — mirror of pde_opt/pde_model.py for the stepping path: `solve` (:68-136) and the
differentiable-rollout objectives `residual_single` / `regularization` / `residuals` / `mse`
(:138-323).  Gradients come from the hand-written adjoint kernel (pde_opt_b200/adjoint.py) through
torch.autograd, the role jax.grad + diffrax adjoints play in the reference.  `train(method="mse")`
and `optimize` (:325-551) are host loops over those gradients (optimistix BFGS in the reference,
torch.optim.LBFGS here: same quasi-Newton family, not the same iterates); `train(method="least_squares")`,
the reference's default, is a Levenberg-Marquardt loop over forward-mode Jacobians from the tangent kernels
(pde_opt_b200/least_squares.py)."""
from typing import Any, Dict

import numpy as np
import torch

from .schedule import constant_step_times
from .solvers import ODETerm
from .utils import check_equation_solver_compatibility, prepare_solver_params


class PDEModel:
    def __init__(self, equation_type, domain, solver_type):
        self.equation_type = equation_type
        self.domain = domain
        self.solver_type = solver_type
        check_equation_solver_compatibility(self.solver_type, self.equation_type)  # pde_model.py:66

    def solve(self, parameters: Dict[str, Any], y0, ts, solver_parameters: Dict[str, Any] = {}, adjoint=None,
              dt0=0.000001, max_steps=1000000, stepsize_controller=None):
        """Integrate from ts[0] to ts[-1] with constant step dt0 and return the solution at `ts`
        (shape (len(ts), *y0.shape)), linearly interpolated inside the step that brackets each
        save time, as diffrax's SaveAt(ts=ts) does with LocalLinearInterpolation
        (pde_model.py:120-136).  `stepsize_controller`: None / ConstantStepSize (the reference's default,
        fused K-step launches) or `pde_opt_b200.stepsize.PIDController` (one launch per attempted step)."""
        equation = self.equation_type(domain=self.domain, **parameters)  # :110
        solver = self.solver_type(**prepare_solver_params(self.solver_type, solver_parameters, equation))  # :112-117
        if stepsize_controller is not None and type(stepsize_controller).__name__ != "ConstantStepSize":
            if not hasattr(stepsize_controller, "adapt"):
                raise NotImplementedError("stepsize_controller must be ConstantStepSize or pde_opt_b200.stepsize.PIDController")
            return self._solve_adaptive(equation, solver, y0, ts, dt0, max_steps, stepsize_controller)
        if type(equation).__name__ == "AdvectionDiffusion2D":
            return self._solve_differentiable(equation, solver, y0, ts, dt0, max_steps, adjoint)
        if self._wants_phasefield_grad(equation, y0):
            return self._solve_differentiable(equation, solver, y0, ts, dt0, max_steps, adjoint)
        if self._wants_closure_grad(equation, y0):
            return self._solve_differentiable(equation, solver, y0, ts, dt0, max_steps, adjoint, unfused=True)
        terms = ODETerm(equation)
        ts = np.asarray([float(t) for t in ts], dtype=np.float32)
        times = constant_step_times(ts[0], ts[-1], dt0, np.float32, max_steps)
        truncated = times[-1] < ts[-1]
        y = y0 if torch.is_tensor(y0) else torch.as_tensor(np.asarray(y0, dtype=np.float32))
        if not y.is_cuda:
            y = y.cuda()
        y = y.to(torch.float32).contiguous()
        out = torch.empty((len(ts),) + tuple(y.shape), dtype=torch.float32, device=y.device)
        i_cur = 0
        for si, s in enumerate(ts):
            j = int(np.searchsorted(times, s, side="left"))
            if j >= len(times):
                out[si:] = float("inf")  # diffrax (throw=False) leaves unreached save slots at inf
                break
            if j == 0 or times[j] == s:
                if j > i_cur:
                    y = solver.rollout(terms, times[i_cur : j + 1], y)
                    i_cur = j
                out[si] = y
                continue
            if j - 1 > i_cur:
                y = solver.rollout(terms, times[i_cur:j], y)
                i_cur = j - 1
            y_b = solver.rollout(terms, times[j - 1 : j + 1], y)
            w = np.float32((s - times[j - 1]) / (times[j] - times[j - 1]))
            out[si] = torch.lerp(y, y_b, float(w))
            y, i_cur = y_b, j
        del truncated
        return out

    def _solve_adaptive(self, equation, solver, y0, ts, dt0, max_steps, controller):
        """diffeqsolve with an adaptive controller: one solver.step per attempt (the GPU step plus the
        `y_error` of solvers.py:61-65), the controller's accept / reject decision on the host, rejected
        steps retried from the same state, `SaveAt(ts)` by linear interpolation inside accepted steps.
        Attempts (accepted or not) count towards `max_steps`; slots not reached stay at inf
        (throw=False, pde_model.py:131)."""
        if not hasattr(solver, "with_error"):
            raise ValueError(f"{type(solver).__name__} provides no error estimate (y_error is None): use ConstantStepSize")
        solver.with_error = True
        terms = ODETerm(equation)
        ts = np.asarray([float(t) for t in ts], dtype=np.float32)
        y = y0 if torch.is_tensor(y0) else torch.as_tensor(np.asarray(y0, dtype=np.float32))
        y = y.to(device="cuda", dtype=torch.float32).contiguous() if not y.is_cuda else y.to(torch.float32).contiguous()
        out = torch.full((len(ts),) + tuple(y.shape), float("inf"), dtype=torch.float32, device=y.device)
        t, t1 = np.float32(ts[0]), np.float32(ts[-1])
        si = 0
        while si < len(ts) and ts[si] <= t:
            out[si] = y
            si += 1
        dt, state, order = float(dt0), controller.init_state(), solver.order(terms)
        self.last_stats = {"accepted": 0, "rejected": 0}
        for _ in range(int(max_steps)):
            if not t < t1:
                break
            tn = np.float32(t + np.float32(dt))
            if tn > t1 - np.float32(1e-6) or tn >= t1:
                tn = t1
            y1, y_err, _, _, _ = solver.step(terms, t, tn, y)
            err = controller.scaled_error(y, y1, y_err)
            keep, dt, state = controller.adapt(float(tn - t), err, order, state)
            if not keep:
                self.last_stats["rejected"] += 1
                continue
            self.last_stats["accepted"] += 1
            while si < len(ts) and ts[si] <= tn:
                out[si] = torch.lerp(y, y1, float(np.float32((ts[si] - t) / (tn - t))))
                si += 1
            y, t = y1, tn
        return out

    # ---- differentiable rollouts (hand-written adjoints) ----------------------------------------------
    @staticmethod
    def _wants_phasefield_grad(equation, y0):
        """Cahn-Hilliard / Allen-Cahn 2-D with tensor-valued closure coefficients (or an initial state
        that requires grad): route through the adjoint-capable rollout."""
        if getattr(equation, "_kind", None) not in ("ch2d", "ac2d", "ch3d") or getattr(equation, "derivs", "") != "fd":
            return False
        if not getattr(equation, "fused", False) or getattr(equation, "control", None) is not None:
            return False
        leaves = equation._mu_c.tensor_leaves() + equation._mob_c.tensor_leaves()
        return any(t.requires_grad for t in leaves) or (torch.is_tensor(y0) and y0.requires_grad)

    @staticmethod
    def _wants_closure_grad(equation, y0):
        """Cahn-Hilliard / Allen-Cahn 2-D (derivs='fd') with a torch closure of the whole field (PeriodicCNN, Mixer2d, ...)
        whose parameters require grad, or with an initial state that does: the unfused differentiable rollout."""
        if getattr(equation, "_kind", None) not in ("ch2d", "ac2d") or getattr(equation, "derivs", "") != "fd":
            return False
        if getattr(equation, "fused", True) or getattr(equation, "control", None) is not None:
            return False
        from .adjoint_nn import closure_parameters

        return bool(closure_parameters(equation)) or (torch.is_tensor(y0) and y0.requires_grad)

    def _solve_differentiable(self, equation, solver, y0, ts, dt0, max_steps, adjoint, unfused=False):
        """solve() on the differentiable paths (advection-diffusion; phase-field equations with tensor
        coefficients): same save-time semantics, every segment an autograd node whose backward is an
        adjoint kernel.  `adjoint` may carry `checkpoint_every` (int): keep every C-th state and recompute in between."""
        from .adjoint import ad_rollout
        from .adjoint_ch import phasefield_rollout

        is_ad = type(equation).__name__ == "AdvectionDiffusion2D"

        ts = np.asarray([float(t) for t in ts], dtype=np.float32)
        times = constant_step_times(ts[0], ts[-1], dt0, np.float32, max_steps)
        y = y0 if torch.is_tensor(y0) else torch.as_tensor(np.asarray(y0, dtype=np.float32))
        y = y.to(device="cuda", dtype=torch.float32) if not y.is_cuda else y.to(torch.float32)
        single = y.dim() == len(self.domain.points)
        if single:
            y = y.unsqueeze(0)
        if is_ad:
            ctrl = equation.control_block(y.shape[0], y.device, self._nseg(equation))
            hold = max(1, -(-(len(times) - 1) // ctrl.shape[1]))
            ck = getattr(adjoint, "checkpoint_every", None)
            A = float(solver.A)

            def roll(y_, seg_times, step0):
                return ad_rollout(equation, y_, ctrl, seg_times, hold=hold, A=A, checkpoint_every=ck, step0=step0)
        elif unfused:
            from .adjoint_nn import given_mu_rollout

            def roll(y_, seg_times, step0):
                return given_mu_rollout(equation, solver, y_, seg_times)
        else:
            ck_pf = getattr(adjoint, "checkpoint_every", None)

            def roll(y_, seg_times, step0):
                return phasefield_rollout(equation, solver, y_, seg_times, checkpoint_every=ck_pf)

        out, i_cur = [], 0
        for s_ in ts:
            j = int(np.searchsorted(times, s_, side="left"))
            if j >= len(times):
                out.append(torch.full_like(y, float("inf")))
                continue
            if j == 0 or times[j] == s_:
                if j > i_cur:
                    y = roll(y, times[i_cur : j + 1], i_cur)
                    i_cur = j
                out.append(y)
                continue
            if j - 1 > i_cur:
                y = roll(y, times[i_cur:j], i_cur)
                i_cur = j - 1
            y_b = roll(y, times[j - 1 : j + 1], j - 1)
            w = float(np.float32((s_ - times[j - 1]) / (times[j] - times[j - 1])))
            out.append(y + (y_b - y) * w)  # LocalLinearInterpolation
            y, i_cur = y_b, j
        ys = torch.stack(out, 0)
        return ys[:, 0] if single else ys

    @staticmethod
    def _nseg(equation):
        v = equation.velocity
        n = 1
        for m in (v.centre[0], v.centre[1], v.p0, v.p1):
            if torch.is_tensor(m) and m.dim() == 2:
                n = max(n, m.shape[1])
        return n

    def residual_single(self, parameters, solver_parameters, y0, values, ts, adjoint=None, dt0=0.000001):
        """values - predicted[1:] for one trajectory (pde_model.py:138-171)."""
        pred = self.solve(parameters, y0, ts, solver_parameters, adjoint=adjoint, dt0=dt0)
        return values - pred[1:]

    def regularization(self, parameters, weights, lambda_reg):
        """lambda * sum_i w_i p_i^2 over the weighted parameters (pde_model.py:173-224).  The reference reduces
        leaf-wise over pytrees; here a parameter may be a tensor / number or a closure object, whose leaves are
        its `tensor_leaves()` (weights: one array for a single leaf, or a sequence with one entry per leaf; None
        entries are skipped)."""
        reg = 0.0
        for key, w in weights.items():
            if w is None:
                continue
            p = parameters[key]
            if hasattr(p, "tensor_leaves"):
                leaves = p.tensor_leaves()
                ws = [w] if (torch.is_tensor(w) or np.isscalar(w) or isinstance(w, np.ndarray)) else list(w)
                if len(ws) != len(leaves):
                    raise ValueError(f"weights[{key!r}] has {len(ws)} entries for {len(leaves)} parameter leaves")
                for wi, leaf in zip(ws, leaves):
                    if wi is not None:
                        reg = reg + lambda_reg * (torch.as_tensor(wi, device=leaf.device) * leaf**2).sum()
            else:
                pt = torch.as_tensor(p)
                reg = reg + lambda_reg * (torch.as_tensor(w, device=pt.device) * pt**2).sum()
        return reg

    def residuals(self, parameters, y0s__values, solver_parameters, ts, weights, lambda_reg, adjoint=None, dt0=0.000001):
        """Batched residuals [batch, timepoints, *shape] and the regularisation term
        (pde_model.py:226-272).  The batch axis is native to the kernels (no vmap)."""
        y0s, values = y0s__values
        pred = self.solve(parameters, y0s, ts, solver_parameters, adjoint=adjoint, dt0=dt0)  # [T, B, ...]
        res = values - pred[1:].transpose(0, 1)
        return res, self.regularization(parameters, weights, lambda_reg)

    def mse(self, parameters, y0s__values, solver_parameters, ts, weights, lambda_reg, adjoint=None, dt0=0.000001):
        """mean(residuals^2) + regularisation (pde_model.py:274-323); a torch scalar whose
        `.backward()` runs the adjoint kernel."""
        res, reg = self.residuals(parameters, y0s__values, solver_parameters, ts, weights, lambda_reg, adjoint, dt0)
        return (res**2).mean() + reg

    # ---- optimisation drivers (host loops over the adjoint gradients) --------------------------------
    @staticmethod
    def _leaves(opt_parameters):
        leaves = []
        for v in opt_parameters.values():
            if torch.is_tensor(v):
                leaves.append(v)
            elif hasattr(v, "tensor_leaves"):
                leaves.extend(v.tensor_leaves())
            elif isinstance(v, torch.nn.Module):  # neural closures: every parameter is a leaf
                leaves.extend(v.parameters())
        leaves = [t for t in leaves if t.is_floating_point()]
        if not leaves:
            raise ValueError("opt_parameters holds no floating-point tensors to optimise")
        for t in leaves:
            t.requires_grad_(True)
        return leaves

    def _minimise(self, loss_fn, leaves, max_steps, verbose=False):
        opt = torch.optim.LBFGS(leaves, lr=1.0, max_iter=int(max_steps), tolerance_grad=1e-10, tolerance_change=1e-14,
                                history_size=20, line_search_fn="strong_wolfe")
        history = []

        def closure():
            opt.zero_grad()
            loss = loss_fn()
            loss.backward()
            history.append(float(loss.detach()))
            if verbose:
                print(f"loss {history[-1]:.6e}")
            return loss

        opt.step(closure)
        self.last_loss_history = history
        return history

    def train(self, data, inds, opt_parameters, other_parameters, solver_parameters, weights, lambda_reg, method="least_squares",
              max_steps=100, dt0=0.000001, verbose=False):
        """Fit `opt_parameters` to observed trajectories (pde_model.py:325-460).  `data = {"ys": [...], "ts":
        [...]}`, `inds` = per trajectory [initial index, later indices...].  method "least_squares" (the
        reference's default): Levenberg-Marquardt over PDEModel.residuals with forward-mode Jacobians (tangent
        kernels; closure coefficients of the finite-difference Cahn-Hilliard / Allen-Cahn equations); method
        "mse": quasi-Newton minimisation of PDEModel.mse over the adjoint gradients (any differentiable
        equation).  Returns the optimised parameters merged with `other_parameters` (tensors updated in place)."""
        if method not in ("least_squares", "mse"):
            raise ValueError(f"unknown method {method!r}: 'least_squares' or 'mse'")
        dev = "cuda"
        y0s = torch.stack([torch.as_tensor(data["ys"][ind[0]], dtype=torch.float32) for ind in inds]).to(dev)
        values = torch.stack([torch.stack([torch.as_tensor(data["ys"][i], dtype=torch.float32) for i in ind[1:]]) for ind in inds]).to(dev)
        ts = np.asarray([data["ts"][i] - data["ts"][inds[0][0]] for i in inds[0]], dtype=np.float32)
        leaves = self._leaves(opt_parameters)
        if method == "least_squares":
            from .functions import Closure
            from .least_squares import levenberg_marquardt

            if not all(isinstance(v, Closure) for v in opt_parameters.values()):
                raise NotImplementedError("method='least_squares' optimises the tensor coefficients of mu / mobility closures "
                                          "(forward-mode tangent kernels); use method='mse' for other parameters")
            self.last_loss_history = levenberg_marquardt(self, {**opt_parameters, **other_parameters}, leaves, y0s, values, ts,
                                                         solver_parameters, weights, lambda_reg, dt0, max_steps, verbose=verbose)
            return {**opt_parameters, **other_parameters}
        self._minimise(lambda: self.mse({**opt_parameters, **other_parameters}, (y0s, values), solver_parameters, ts, weights,
                                        lambda_reg, dt0=dt0), leaves, max_steps, verbose)
        return {**opt_parameters, **other_parameters}

    def optimize(self, objective_function, y0, ts, opt_parameters, other_parameters, solver_parameters, weights, lambda_reg,
                 max_steps=100, dt0=0.000001, verbose=False):
        """Minimise objective_function(solution) + regularisation over `opt_parameters`
        (pde_model.py:462-551); `solution` has shape (len(ts), *y0.shape)."""
        leaves = self._leaves(opt_parameters)

        def loss_fn():
            allp = {**opt_parameters, **other_parameters}
            sol = self.solve(allp, y0, ts, solver_parameters, dt0=dt0)
            return objective_function(sol) + self.regularization(allp, weights, lambda_reg)

        self._minimise(loss_fn, leaves, max_steps, verbose)
        return {**opt_parameters, **other_parameters}
