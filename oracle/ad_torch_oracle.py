"""Gradient oracle: torch-CPU twin of the advection-diffusion rollout + torch.autograd.

TEST INFRASTRUCTURE ONLY (see oracle/pde_oracle.py).  The reference differentiates rollouts with
jax.grad through diffrax (pde_model.py:226-323); JAX cannot be installed here, so torch.autograd
on a float64 restatement of the same arithmetic stands in for jax.grad.  The arithmetic restates
  * AdvectionDiffusion2D.rhs / gaussian_velocity of oracle/pde_oracle.py (recovered equation,
    notebooks/run_advection_diffusion.ipynb cells 0-2), and
  * SemiImplicitFourierSpectral.step (pde_opt/numerics/solvers.py:56-70),
and is checked against that NumPy oracle and against central finite differences in
tests/test_oracle_kat.py.  "Parity unpinned" by the reference itself: it holds no gradient tests."""
import math

import numpy as np
import torch


def _mesh(points, box, dtype):
    axes = []
    for (lo, hi), n in zip(box, points):
        h = (hi - lo) / n
        axes.append(torch.linspace(lo + h / 2, hi - h / 2, n, dtype=dtype))  # domains.py:36-42
    return torch.meshgrid(*axes, indexing="ij")  # domains.py:54-56


def _kmesh(points, box, dtype):
    ks = []
    for (lo, hi), n in zip(box, points):
        h = (hi - lo) / n
        ks.append(torch.fft.fftfreq(n, h, dtype=dtype))  # domains.py:44-47
    return torch.meshgrid(*ks, indexing="ij")


def rollout(y0, ctrl, dts, points, box, D, A=1.0, hold=None):
    """y0 [B,nx,ny]; ctrl [B,nseg,4] = (cx, cy, p0, p1); dts: sequence of step lengths.
    Differentiable w.r.t. y0 and ctrl.  Returns the final state."""
    dtype = y0.dtype
    cdtype = torch.complex128 if dtype == torch.float64 else torch.complex64
    xs, ys = _mesh(points, box, dtype)
    kx, ky = _kmesh(points, box, dtype)
    ikx = (2j * math.pi * kx).to(cdtype)
    iky = (2j * math.pi * ky).to(cdtype)
    k2 = ikx**2 + iky**2
    sigma = -D * k2  # our IMEX symbol (oracle/pde_oracle.py AdvectionDiffusion2D)
    nseg = ctrl.shape[1]
    K = len(dts)
    if hold is None:
        hold = max(1, -(-K // nseg))
    y = y0
    for k in range(K):
        s = min(k // hold, nseg - 1)
        cx, cy, p0, p1 = (ctrl[:, s, j][:, None, None] for j in range(4))
        e = torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2.0 * p1))
        vx = p0 * (-(xs - cx) / p1 * e)
        vy = p0 * (-(ys - cy) / p1 * e)
        uh = torch.fft.fftn(y, dim=(-2, -1))
        fx = torch.fft.fftn(vx * y, dim=(-2, -1))
        fy = torch.fft.fftn(vy * y, dim=(-2, -1))
        spec = -(ikx * fx + iky * fy) + D * k2 * uh
        f0 = torch.fft.ifftn(spec, dim=(-2, -1)).real
        dt = float(dts[k])
        g = torch.fft.ifftn(torch.fft.fftn(f0, dim=(-2, -1)) / (1.0 + A * dt * sigma), dim=(-2, -1)).real  # solvers.py:62-63
        y = y + dt * g
    return y
