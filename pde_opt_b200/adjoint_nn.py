"""Differentiable rollout for phase-field equations whose mu (and / or mobility) is a closure of the whole field — the
reference's neural closures PeriodicCNN / Mixer2d (docs/notebooks/optimization_neural_network.ipynb trains such a mu
by jax.grad through diffeqsolve, pde_model.py:274-323) or any torch callable.

Forward: per step the closure runs in torch on the whole batch, the stencils in pdeopt_rhs_given_mu_batched and the
spectral filter in pdeopt_sifs_filter_batched; the state before every step is kept.  Backward, per step in reverse:
pdeopt_phasefield_adjoint_given_mu returns the cotangents of mu_h and of the mobility (mubar, dbar) and
lam1 - kappa lap(mubar); torch back-propagates (mubar, dbar) through the closure, which yields the closure's share of the
state cotangent and the gradients of its parameters.  The CUDA side never sees the network."""
import ctypes

import numpy as np
import torch

from . import _lib


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()


def closure_parameters(eq):
    """Trainable torch parameters of the equation's closures (mu, D / R), in a fixed order."""
    ps = []
    for name in ("mu", "D", "R"):
        fn = getattr(eq, name, None)
        if isinstance(fn, torch.nn.Module):
            ps += [p for p in fn.parameters() if p.requires_grad]
    return ps


class _GivenMuRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, eq, solver, dts, sym, *params):
        y = y0.contiguous()
        plan = solver.filter_plan(tuple(y.shape[-2:]))
        traj = torch.empty((len(dts),) + tuple(y.shape), dtype=torch.float32, device=y.device)
        for k, dt in enumerate(dts):
            traj[k].copy_(y)
            f0 = eq._rhs_given_mu(y)
            y = plan.filter(y, f0, dt, sym)
        ctx.eq, ctx.plan, ctx.dts, ctx.sym, ctx.nparams = eq, plan, dts, sym, len(params)
        ctx.save_for_backward(traj, *params)
        return y

    @staticmethod
    def backward(ctx, gy):
        traj, *params = ctx.saved_tensors
        eq, dts, sym = ctx.eq, ctx.dts, ctx.sym
        lib = _lib.load()
        gm = eq._gm_plan  # built by the forward's _rhs_given_mu
        mob_fn = eq.D if eq._kind == "ch2d" else eq.R
        lam = gy.contiguous().clone()
        B = lam.shape[0]
        n = lam.numel()
        work = torch.empty(4 * n, dtype=torch.float32, device=lam.device)
        mubar, dbar, base = torch.empty_like(lam), torch.empty_like(lam), torch.empty_like(lam)
        gparams = [torch.zeros_like(p) for p in params]
        for k in range(len(dts) - 1, -1, -1):
            u = traj[k].detach().clone().requires_grad_(True)
            with torch.enable_grad():
                muh = eq.mu(u).to(torch.float32)
                mobv = mob_fn(u)
                mobv = mobv.to(torch.float32) if torch.is_tensor(mobv) else torch.full_like(u, float(mobv))
            with _lib.device_of(lam):
                _lib.check(lib.pdeopt_phasefield_adjoint_given_mu(gm._h, _vp(u.detach()), _vp(muh.detach().contiguous()),
                                                                  _vp(mobv.detach().contiguous()), _vp(lam), _vp(base), _vp(mubar), _vp(dbar),
                                                                  B, float(dts[k]), _vp(sym), _vp(work), _lib.stream_ptr(lam)))
            outs, couts = [muh], [mubar]
            if mobv.requires_grad:
                outs.append(mobv)
                couts.append(dbar)
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):  # float32 weight gradients, not TF32
                grads = torch.autograd.grad(outs, [u] + list(params), couts, allow_unused=True)
            lam = base + (grads[0] if grads[0] is not None else 0.0)
            for g_acc, g in zip(gparams, grads[1:]):
                if g is not None:
                    g_acc += g
        return (lam, None, None, None, None, *gparams)


def given_mu_rollout(eq, solver, y0, times):
    """y after the steps between consecutive `times`, differentiable w.r.t. y0 and the parameters of the equation's torch
    closures.  y0: [B, nx, ny] float32 CUDA."""
    times = np.asarray(times, dtype=np.float32)
    dts = np.ascontiguousarray((times[1:] - times[:-1]).astype(np.float32))
    return _GivenMuRollout.apply(y0, eq, solver, dts, solver.symbol_on(y0.device), *closure_parameters(eq))
