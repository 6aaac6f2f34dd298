"""Differentiable advection-diffusion rollout: forward kernel + hand-written adjoint kernel wrapped
in a torch.autograd.Function.

This is the role `jax.custom_vjp` plays in the north star (JAX is not installable in this image,
SURVEY F3/F5): the forward launches pdeopt_ad_rollout_fwd (saving the state at the start of every
step), the backward launches pdeopt_ad_rollout_bwd, the discrete adjoint of exactly those steps.
It replaces reverse-mode differentiation through `diffrax.diffeqsolve` with
RecursiveCheckpointAdjoint as used by PDEModel.residuals / mse (pde_model.py:226-323).

Gradients flow to the initial state `y0` and to the control block `ctrl[B, nseg, 4]` =
(cx, cy, p0, p1) per control segment."""
import ctypes

import numpy as np
import torch

from . import _lib


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()


def _fwd(desc, y_in, y_out, dts, tables, ctrl, hold, step0, traj):
    lib = _lib.load()
    stream = ctypes.c_void_p(torch.cuda.current_stream(y_in.device).cuda_stream)
    B, nseg = ctrl.shape[0], ctrl.shape[1]
    done, src = 0, y_in
    while done < len(dts):
        k = min(_lib.MAX_FUSED_STEPS, len(dts) - done)
        tr = traj[done:] if traj is not None else None
        st = lib.pdeopt_ad_rollout_fwd(
            ctypes.byref(desc), _vp(src), _vp(y_out), B, k, dts[done:].ctypes.data_as(ctypes.c_void_p), _vp(tables),
            _vp(ctrl), nseg, hold, step0 + done, _vp(tr), tr.stride(0) if tr is not None else 0, stream,
        )
        _lib.check(st)
        src = y_out
        done += k


def _bwd(desc, traj, lam, dts, tables, ctrl, hold, step0, gctrl):
    """In place on `lam`: cotangent after the len(dts) steps -> cotangent before them."""
    lib = _lib.load()
    stream = ctypes.c_void_p(torch.cuda.current_stream(lam.device).cuda_stream)
    B, nseg = ctrl.shape[0], ctrl.shape[1]
    end = len(dts)
    while end > 0:
        k = min(_lib.MAX_FUSED_STEPS, end)
        beg = end - k
        tr = traj[beg:]
        st = lib.pdeopt_ad_rollout_bwd(
            ctypes.byref(desc), _vp(tr), tr.stride(0), _vp(lam), _vp(lam), B, k,
            dts[beg:].ctypes.data_as(ctypes.c_void_p), _vp(tables), _vp(ctrl), nseg, hold, step0 + beg, _vp(gctrl), stream,
        )
        _lib.check(st)
        end = beg


class _ADRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, ctrl, eq, dts, tables, hold, checkpoint_every, step0=0):
        desc = eq.ad_desc()
        y0c, ctrlc = y0.contiguous(), ctrl.contiguous()
        K = len(dts)
        need_grad = y0.requires_grad or ctrl.requires_grad
        y1 = torch.empty_like(y0c)
        ctx.desc, ctx.dts, ctx.tables, ctx.hold, ctx.ckpt, ctx.step0 = desc, dts, tables, hold, checkpoint_every, int(step0)
        step0 = int(step0)
        if not need_grad:
            _fwd(desc, y0c, y1, dts, tables, ctrlc, hold, step0, None)
            return y1
        if checkpoint_every is None:
            # the whole trajectory lives in HBM (500 steps x 512 envs x 64 KB = 16 GiB of 180 GB)
            traj = torch.empty((K, 2 * ((y0c.shape[0] + 1) // 2)) + tuple(y0c.shape[1:]), dtype=torch.float32, device=y0c.device)
            _fwd(desc, y0c, y1, dts, tables, ctrlc, hold, step0, traj)
            ctx.save_for_backward(ctrlc, traj)
        else:
            S = int(checkpoint_every)
            cps = []
            y = y0c
            for beg in range(0, K, S):
                cps.append(y if beg == 0 else y.clone())
                nxt = torch.empty_like(y0c)
                _fwd(desc, y, nxt, dts[beg : beg + S], tables, ctrlc, hold, step0 + beg, None)
                y = nxt
            y1 = y
            ctx.save_for_backward(ctrlc, *cps)
        return y1

    @staticmethod
    def backward(ctx, gy1):
        desc, dts, tables, hold, S, step0 = ctx.desc, ctx.dts, ctx.tables, ctx.hold, ctx.ckpt, ctx.step0
        ctrl = ctx.saved_tensors[0]
        K = len(dts)
        lam = gy1.contiguous().clone()
        gctrl = torch.zeros_like(ctrl)
        if S is None:
            traj = ctx.saved_tensors[1]
            _bwd(desc, traj, lam, dts, tables, ctrl, hold, step0, gctrl)
        else:
            cps = ctx.saved_tensors[1:]
            seg = torch.empty((min(S, K), 2 * ((lam.shape[0] + 1) // 2)) + tuple(lam.shape[1:]), dtype=torch.float32, device=lam.device)
            scratch = torch.empty_like(lam)
            begs = list(range(0, K, S))
            for ci in reversed(range(len(begs))):
                beg = begs[ci]
                d = dts[beg : beg + S]
                _fwd(desc, cps[ci], scratch, d, tables, ctrl, hold, step0 + beg, seg)  # recompute the segment
                _bwd(desc, seg, lam, d, tables, ctrl, hold, step0 + beg, gctrl)
        return lam, gctrl, None, None, None, None, None, None


def ad_rollout(eq, y0, ctrl, times, hold=None, A=1.0, checkpoint_every=None, step0=0):
    """Differentiable rollout of AdvectionDiffusion2D over the step boundaries `times`.

    y0   : [B, nx, ny] float32 CUDA
    ctrl : [B, nseg, 4] float32 CUDA (cx, cy, p0, p1); segment s is held for `hold` numeric steps
           (default: the whole rollout divided evenly over nseg)
    step0: index of the first step of `times` within the whole rollout (selects the control segment
           when a rollout is issued in several calls)
    Returns the final state [B, nx, ny]; differentiable w.r.t. y0 and ctrl."""
    times = np.asarray(times, dtype=np.float32)
    dts = np.ascontiguousarray((times[1:] - times[:-1]).astype(np.float32))
    assert y0.is_cuda and y0.dtype == torch.float32 and y0.dim() == 3
    assert ctrl.is_cuda and ctrl.dtype == torch.float32 and ctrl.dim() == 3 and ctrl.shape[2] == 4
    assert ctrl.shape[0] == y0.shape[0]
    if hold is None:
        hold = max(1, -(-len(dts) // ctrl.shape[1]))
    tables = eq.tables_on(y0.device, A)
    return _ADRollout.apply(y0, ctrl, eq, dts, tables, int(hold), checkpoint_every, int(step0))
