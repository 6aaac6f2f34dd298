"""Constant-step time grid of diffrax.diffeqsolve(..., ConstantStepSize()) on the host.

The reference delegates the step loop to diffrax (call sites pde_env.py:293-303,
pde_model.py:120-134).  The fused kernels take explicit per-step lengths dt[k] = t1 - t0
(solvers.py:58), so the loop's floating-point time accumulation and end clipping are
reproduced here: tnext <- tprev + (tnext - tprev) in the working precision, and
tnext := t1 when tnext > t1 - tol (tol 1e-6 in float32, 1e-10 in float64; diffrax
`_clip_to_end`)."""
import numpy as np


def constant_step_times(t0, t1, dt0, dtype=np.float32, max_steps=1_000_000):
    tt = np.dtype(dtype).type
    tol = tt(1e-6) if np.dtype(dtype) == np.float32 else tt(1e-10)
    t0, t1, dt0 = tt(t0), tt(t1), tt(dt0)
    out = [t0]
    tprev, tnext = t0, tt(t0 + dt0)
    if tnext > t1 - tol:
        tnext = t1
    n = 0
    while tprev < t1 and n < max_steps:
        out.append(tnext)
        step = tt(tnext - tprev)
        tprev = tnext
        tnext = tt(tprev + step)
        if tnext > t1 - tol:
            tnext = t1
        n += 1
    return np.asarray(out, dtype=dtype)


def step_lengths(times):
    """dt[k] = t[k+1] - t[k] evaluated in the working precision (solvers.py:58)."""
    times = np.asarray(times)
    return (times[1:] - times[:-1]).astype(times.dtype)
