"""Gradient oracle for the phase-field equations: torch-CPU float64 twin of
CahnHilliard2DPeriodic.rhs_fd (pde_opt/numerics/equations/cahn_hilliard.py:89-109),
AllenCahn2DPeriodic.rhs_fd (allen_cahn.py:81-84), the roll stencils
(pde_opt/numerics/utils/derivatives.py:8-66), the Legendre closures
(pde_opt/numerics/functions/legendre.py:12-74) and SemiImplicitFourierSpectral.step
(pde_opt/numerics/solvers.py:56-70), differentiated by torch.autograd (stand-in for jax.grad).

TEST INFRASTRUCTURE ONLY.  Checked against the NumPy oracle in tests/test_oracle_kat.py."""
import math

import torch


def legendre(params, x):  # legendre.py:19-34
    result = params[0] * torch.ones_like(x)
    if len(params) > 1:
        result = result + params[1] * x
    p_prev, p_curr = torch.ones_like(x), x
    for n in range(2, len(params)):
        p_next = ((2 * n - 1) * x * p_curr - (n - 1) * p_prev) / n
        result = result + params[n] * p_next
        p_prev, p_curr = p_curr, p_next
    return result


def mu_legendre(params, c, log_prior=False):  # legendre.py:56-74
    r = legendre(params, 2.0 * c - 1.0)
    return r + torch.log(c / (1.0 - c)) if log_prior else r


def D_legendre(params, c):  # legendre.py:37-53
    return torch.exp(legendre(params, 2.0 * c - 1.0))


def _axes(h):
    return list(zip(range(-len(h), 0), h))


def _lap(u, h):  # derivatives.py:8-21 (2-D and 3-D)
    return sum((torch.roll(u, -1, ax) - 2 * u + torch.roll(u, 1, ax)) / hh**2 for ax, hh in _axes(h))


def rhs_ch(u, h, kappa, mu_fn, D_fn):  # cahn_hilliard.py:89-109, :177-200
    mu = mu_fn(u) - kappa * _lap(u, h)
    D = D_fn(u)
    out = 0
    for ax, hh in _axes(h):
        F = 0.5 * (D + torch.roll(D, -1, ax)) * ((torch.roll(mu, -1, ax) - mu) / hh)
        out = out + (F - torch.roll(F, 1, ax)) / hh
    return out


def rhs_ac(u, h, kappa, mu_fn, R_fn):  # allen_cahn.py:81-84
    return -R_fn(u) * (mu_fn(u) - kappa * _lap(u, h))


def rollout(y0, dts, points, box, kappa, A, mu_fn, D_fn, kind="ch"):
    """y0 [B, nx, ny] or [B, nx, ny, nz]; returns the state after len(dts) semi-implicit steps
    (differentiable)."""
    dtype = y0.dtype
    h = [(hi - lo) / n for (lo, hi), n in zip(box, points)]
    ks = torch.meshgrid(*[torch.fft.fftfreq(n, hh, dtype=dtype) for n, hh in zip(points, h)], indexing="ij")
    k2 = sum((2j * math.pi * k) ** 2 for k in ks)
    sigma = kappa * k2**2 if kind == "ch" else -kappa * k2
    y = y0
    for dt in dts:
        dt = float(dt)
        f0 = rhs_ch(y, h, kappa, mu_fn, D_fn) if kind == "ch" else rhs_ac(y, h, kappa, mu_fn, D_fn)
        dims = tuple(range(-len(points), 0))
        g = torch.fft.ifftn(torch.fft.fftn(f0, dim=dims) / (1.0 + A * dt * sigma), dim=dims).real
        y = y + dt * g
    return y
