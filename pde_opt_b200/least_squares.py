"""Levenberg-Marquardt over PDEModel.residuals with forward-mode Jacobians — the reference's default training
method (pde_model.py:334,404-428: optimistix.LevenbergMarquardt(rtol=1e-8, atol=1e-8) on `residuals` with
diffrax's ForwardMode adjoint).

The Jacobian of the residuals w.r.t. the optimised closure coefficients comes from the tangent kernels
(pdeopt_phasefield_tangent_steps): one tangent field per coefficient is pushed through every semi-implicit step
next to the forward rollout (pdeopt_sifs_rollout_fwd keeps the step-start states), all directions in one launch.
Only the small (ndir x ndir) normal equations are formed and solved on the host side of torch.

optimistix is not installable here; its LevenbergMarquardt (a damped Gauss-Newton step
(J^T J + lambda I) delta = -J^T r whose damping 1/step_size follows a classical trust-region rule: ratio of actual
to predicted reduction >= 0.99 -> step_size * 3.5, <= 0.01 -> reject and step_size * 0.25) is restated from memory
of the optimistix source: the iterates are NOT pinned against the reference's."""
import numpy as np
import torch

from . import _lib
from .schedule import constant_step_times
from .utils import prepare_solver_params


def _direction_table(equation, leaves):
    """One tangent direction per scalar entry of every optimised leaf: rows of (dmu, dmob) [ndir, 16]."""
    rows = []
    for leaf in leaves:
        if leaf is getattr(equation._mu_c, "coef", None):
            rows += [("mu", i) for i in range(leaf.numel())]
        elif leaf is getattr(equation._mob_c, "coef", None):
            rows += [("mob", i) for i in range(leaf.numel())]
        else:
            raise NotImplementedError("method='least_squares' differentiates the tensor coefficients of the mu / mobility closures of the "
                                      "finite-difference Cahn-Hilliard / Allen-Cahn equations; use method='mse' for other leaves")
    dmu = torch.zeros((len(rows), _lib.MAX_COEF), dtype=torch.float32)
    dmob = torch.zeros_like(dmu)
    for r, (kind, i) in enumerate(rows):
        (dmu if kind == "mu" else dmob)[r, i] = 1.0
    return dmu, dmob


def predictions_and_tangents(model, parameters, leaves, y0s, ts, solver_parameters, dt0, max_steps=1_000_000, segment=256):
    """pred [T, B, nx, ny] exactly as PDEModel.solve returns it (constant steps, SaveAt(ts) by linear interpolation
    inside the bracketing step) and dpred [ndir, T, B, nx, ny] = d pred / d theta_dir."""
    equation = model.equation_type(domain=model.domain, **parameters)
    solver = model.solver_type(**prepare_solver_params(model.solver_type, solver_parameters, equation))
    if getattr(equation, "_kind", None) not in ("ch2d", "ac2d") or equation.derivs != "fd" or not equation.fused \
            or getattr(equation, "control", None) is not None:
        raise NotImplementedError("method='least_squares' needs a 2-D Cahn-Hilliard / Allen-Cahn equation with derivs='fd', "
                                  "enumerated closures and no control forcing; use method='mse'")
    plan = equation.plan()
    dev = y0s.device
    sym = solver.symbol_on(dev)
    dmu, dmob = _direction_table(equation, leaves)
    dmu, dmob = dmu.to(dev), dmob.to(dev)
    ndir = dmu.shape[0]
    ts = np.asarray([float(t) for t in ts], dtype=np.float32)
    times = constant_step_times(ts[0], ts[-1], dt0, np.float32, max_steps)
    dts = np.ascontiguousarray((times[1:] - times[:-1]).astype(np.float32))
    K = len(dts)
    # step-boundary indices at which the state (and its tangents) are needed
    plan_pts = []  # per save time: (j0, j1, w): value = lerp(state[j0], state[j1], w)
    for s in ts:
        j = int(np.searchsorted(times, s, side="left"))
        if j >= len(times):
            plan_pts.append(None)
        elif j == 0 or times[j] == s:
            plan_pts.append((j, j, 0.0))
        else:
            plan_pts.append((j - 1, j, float(np.float32((s - times[j - 1]) / (times[j] - times[j - 1])))))
    need = sorted({j for p in plan_pts if p for j in p[:2]})
    y = y0s.to(torch.float32).contiguous()
    v = torch.zeros((ndir,) + tuple(y.shape), dtype=torch.float32, device=dev)
    states, tangents = {}, {}
    k = 0
    if 0 in need:
        states[0], tangents[0] = y.clone(), v.clone()
    stops = [j for j in need if j > 0] or [0]
    for stop in stops:
        while k < stop:
            n = min(segment, stop - k)
            y_next, traj = plan.rollout_fwd(y, dts[k:k + n], sym, save_every=1)
            plan.tangent_steps(traj, v, dts[k:k + n], dmu, dmob, sym)
            y, k = y_next, k + n
        states[stop], tangents[stop] = y.clone(), v.clone()
    del K
    pred = torch.full((len(ts),) + tuple(y.shape), float("inf"), dtype=torch.float32, device=dev)
    dpred = torch.zeros((ndir, len(ts)) + tuple(y.shape), dtype=torch.float32, device=dev)
    for si, p in enumerate(plan_pts):
        if p is None:
            continue  # diffrax (throw=False) leaves unreached save slots at inf
        j0, j1, w = p
        pred[si] = states[j0] if j0 == j1 else torch.lerp(states[j0], states[j1], w)
        dpred[:, si] = tangents[j0] if j0 == j1 else torch.lerp(tangents[j0], tangents[j1], w)
    return pred, dpred


def residuals_and_jacobian(model, parameters, leaves, y0s, values, ts, solver_parameters, weights, lambda_reg, dt0):
    """r = (values - pred[1:], reg) flattened as optimistix flattens the pytree PDEModel.residuals returns
    (pde_model.py:226-272), and J = dr / dtheta [len(r), ndir] (float32 on the device; the regularisation row last)."""
    pred, dpred = predictions_and_tangents(model, parameters, leaves, y0s, ts, solver_parameters, dt0)
    res = values - pred[1:].transpose(0, 1)                    # [B, T-1, ...]
    J = -dpred[:, 1:].transpose(1, 2).reshape(dpred.shape[0], -1)  # [ndir, B*(T-1)*n]
    reg = model.regularization(parameters, weights, lambda_reg)
    theta = torch.cat([t.detach().reshape(-1) for t in leaves]).to(torch.float64)
    dreg = torch.zeros_like(theta)
    if torch.is_tensor(reg) and reg.requires_grad:
        g = torch.autograd.grad(reg, leaves, allow_unused=True)
        dreg = torch.cat([(gi if gi is not None else torch.zeros_like(t)).reshape(-1) for gi, t in zip(g, leaves)]).to(torch.float64)
    return res.reshape(-1), J, float(reg.detach()) if torch.is_tensor(reg) else float(reg), dreg, theta


def lm_iterate(evaluate, theta0, max_steps=100, rtol=1e-8, atol=1e-8, verbose=False):
    """The Levenberg-Marquardt iteration itself, independent of where residuals and Jacobians come from (host logic,
    unit-tested on the CPU): `evaluate(theta) -> (f, JtJ, Jtr)` with f = 1/2 |r|^2, float64 CPU tensors.  Damped
    Gauss-Newton step (J^T J + I / step_size) delta = -J^T r; classical trust-region update of step_size from the ratio of
    actual to predicted reduction (accept above 0.01, grow x 3.5 above 0.99, shrink x 0.25 on rejection); Cauchy
    termination on both the step and the loss.  Returns (theta, loss history)."""
    theta = theta0.clone()
    f, JtJ, Jtr = evaluate(theta)
    history = [f]
    step_size = 1.0
    for it in range(int(max_steps)):
        lam = 1.0 / step_size
        delta = -torch.linalg.solve(JtJ + lam * torch.eye(JtJ.shape[0], dtype=torch.float64), Jtr)
        predicted = float(Jtr @ delta + 0.5 * delta @ (JtJ @ delta))  # model reduction (negative)
        f_new, JtJ_new, Jtr_new = evaluate(theta + delta)
        ratio = (f_new - f) / predicted if predicted < 0 else -1.0
        accept = bool(np.isfinite(f_new)) and ratio > 0.01
        if verbose:
            print(f"LM step {it}: loss {f:.6e} -> {f_new:.6e}  step_size {step_size:.3g}  {'accepted' if accept else 'rejected'}")
        if accept:
            small_y = bool((delta.abs() <= atol + rtol * theta.abs()).all())
            small_f = abs(f_new - f) <= atol + rtol * abs(f)
            theta, f, JtJ, Jtr = theta + delta, f_new, JtJ_new, Jtr_new
            history.append(f)
            if ratio >= 0.99:
                step_size *= 3.5
            if small_y and small_f:
                break
        else:
            step_size *= 0.25
            if step_size < 1e-30:
                break
    return theta, history


def levenberg_marquardt(model, parameters, leaves, y0s, values, ts, solver_parameters, weights, lambda_reg, dt0, max_steps=100,
                        rtol=1e-8, atol=1e-8, verbose=False):
    """Minimise 1/2 (|values - pred[1:]|^2 + reg^2) over the entries of `leaves` (updated in place)."""

    def assign(theta):
        o = 0
        with torch.no_grad():
            for t in leaves:
                t.copy_(theta[o:o + t.numel()].reshape(t.shape).to(t.dtype))
                o += t.numel()

    def evaluate(theta):
        assign(theta)
        with torch.enable_grad():
            r, J, reg, dreg, _ = residuals_and_jacobian(model, parameters, leaves, y0s, values, ts, solver_parameters, weights,
                                                        lambda_reg, dt0)
        f = 0.5 * (float((r.double() ** 2).sum()) + reg**2)
        Jd = J.double()
        JtJ = Jd @ Jd.T + torch.outer(dreg, dreg).to(Jd.device)
        Jtr = Jd @ r.double() + (dreg * reg).to(Jd.device)
        return f, JtJ.cpu(), Jtr.cpu()

    theta0 = torch.cat([t.detach().reshape(-1) for t in leaves]).to(torch.float64).cpu()
    theta, history = lm_iterate(evaluate, theta0, max_steps, rtol, atol, verbose)
    assign(theta)
    return history
