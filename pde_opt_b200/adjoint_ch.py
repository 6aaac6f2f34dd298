"""Differentiable rollout of the finite-difference Cahn-Hilliard / Allen-Cahn equations: the fused
forward kernels (2-D: one launch per 512 steps that also writes the state before every step to HBM) and the adjoint step
(pdeopt_phasefield_adjoint_step) behind a torch.autograd.Function — the custom_vjp of the north star
for the reference's main training use case, fitting the coefficients of mu and D
(docs/notebooks/optimization_3D.ipynb; pde_model.py:226-460).

Gradients flow to the initial state and to the closure coefficients (`mu` / `D` closures built from
1-D torch tensors, shared by the batch)."""
import numpy as np
import torch

from . import _lib


class _PhaseFieldRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, mu_coef, mob_coef, eq, dts, sym, checkpoint_every):
        plan = eq.plan()  # built from the current coefficient values
        ctx.is3d = getattr(eq, "_kind", None) == "ch3d"
        y = y0.contiguous()
        K = len(dts)
        need_grad = y0.requires_grad or (mu_coef is not None and mu_coef.requires_grad) or (mob_coef is not None and mob_coef.requires_grad)
        if not need_grad:
            return plan.step(y, dts, sym)
        ctx.ck = None
        if ctx.is3d:
            traj = torch.empty((K + 1,) + tuple(y.shape), dtype=torch.float32, device=y.device)
            traj[0].copy_(y)
            for k in range(K):
                plan.step(traj[k], dts[k : k + 1], sym, out=traj[k + 1])
            y_end = traj[K].clone()
        else:
            # one fused launch per 512 steps; the kernel writes the kept states itself (pdeopt_sifs_rollout_fwd):
            # every step-start state, or every `checkpoint_every`-th one (the segments in between are re-run
            # during the backward sweep: K / C + C states in HBM instead of K)
            C = int(checkpoint_every) if checkpoint_every else 1
            if C > 1 and C < K:
                ctx.ck = min(C, _lib.MAX_FUSED_STEPS)
            y_end, traj = plan.rollout_fwd(y, dts, sym, save_every=ctx.ck or 1)
        ctx.plan, ctx.dts, ctx.sym = plan, dts, sym
        ctx.has_mu, ctx.has_mob = mu_coef is not None, mob_coef is not None
        ctx.n_mu = int(mu_coef.numel()) if mu_coef is not None else 0
        ctx.n_mob = int(mob_coef.numel()) if mob_coef is not None else 0
        ctx.save_for_backward(traj)
        return y_end

    @staticmethod
    def backward(ctx, gy):
        (traj,) = ctx.saved_tensors
        plan, dts, sym = ctx.plan, ctx.dts, ctx.sym
        lam = gy.contiguous().clone()
        B = lam.shape[0]
        gmu = torch.zeros((B, _lib.MAX_COEF), dtype=torch.float64, device=lam.device)
        gmob = torch.zeros_like(gmu)
        if ctx.is3d:
            for k in range(len(dts) - 1, -1, -1):
                plan.adjoint_step(traj[k], lam, dts[k], sym, gmu, gmob)
        elif ctx.ck is None:
            plan.rollout_bwd(traj, lam, dts, sym, gmu, gmob)  # fused: the cotangent stays on chip for 512 steps
        else:
            C, K = ctx.ck, len(dts)
            for s in range(traj.shape[0] - 1, -1, -1):
                k0, k1 = s * C, min(K, (s + 1) * C)
                _, seg = plan.rollout_fwd(traj[s], dts[k0:k1], sym, save_every=1)  # recompute the segment's states
                plan.rollout_bwd(seg, lam, dts[k0:k1], sym, gmu, gmob)
        g_mu = gmu.sum(0)[: ctx.n_mu].to(torch.float32) if ctx.has_mu else None
        g_mob = gmob.sum(0)[: ctx.n_mob].to(torch.float32) if ctx.has_mob else None
        return lam, g_mu, g_mob, None, None, None, None


def phasefield_rollout(eq, solver, y0, times, checkpoint_every=None):
    """Differentiable rollout of CahnHilliard2DPeriodic / AllenCahn2DPeriodic / CahnHilliard3DPeriodic
    (derivs='fd', enumerated closures, no control forcing) over the step boundaries `times`.
    y0: [B, nx, ny] or [B, nx, ny, nz] float32 CUDA.
    Differentiable w.r.t. y0 and the tensor coefficients of eq.mu and of the mobility closure.
    `checkpoint_every=C` (2-D): keep every C-th state only and re-run the segments during the backward sweep."""
    if not getattr(eq, "fused", False) or eq.derivs != "fd" or getattr(eq, "control", None) is not None:
        raise NotImplementedError("the adjoint needs derivs='fd', enumerated closures and no control forcing")
    times = np.asarray(times, dtype=np.float32)
    dts = np.ascontiguousarray((times[1:] - times[:-1]).astype(np.float32))
    mu_c, mob_c = eq._mu_c, eq._mob_c
    mu_t = mu_c.coef if mu_c.tensor_leaves() else None
    mob_t = mob_c.coef if mob_c.tensor_leaves() else None
    sym = solver.symbol_pos_on(y0.device) if getattr(eq, "_kind", None) == "ch3d" else solver.symbol_on(y0.device)
    return _PhaseFieldRollout.apply(y0, mu_t, mob_t, eq, dts, sym, checkpoint_every)
