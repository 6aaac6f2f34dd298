"""Smoothed-boundary Cahn-Hilliard and Allen-Cahn equations (arbitrary geometries through the level set
`domain.geometry.smooth`).  Mirror of pde_opt/numerics/equations/cahn_hilliard.py:203-289 and allen_cahn.py:87-159:
same dataclass fields, same `left_half` masks, `rhs(state, t)` on CUDA tensors.  The closures f, mu, D / R are
arbitrary callables (evaluated here on the whole batch, torch tensors in / out); the stencils run in
pdeopt_sbm_rhs_batched.  These equations carry no `fourier_symbol`: like the reference they are integrated with an
explicit solver (`pde_opt_b200.solvers.Dopri5`, `Euler`), not with SemiImplicitFourierSpectral."""
import ctypes
import dataclasses
import math
from typing import Any

import numpy as np

from ..domains import Domain
from .base_eq import BaseEquation


class _SmoothedBoundary2D(BaseEquation):
    _sbm_kind = 0

    def _setup(self, side):
        if self.derivs != "fd":
            raise ValueError(f"Invalid derivative type: {self.derivs}")
        if self.domain.geometry is None:
            raise ValueError("smoothed-boundary equations need domain.geometry (a Shape)")
        psi = np.asarray(self.domain.geometry.smooth, dtype=np.float64)
        self.hx, self.hy = self.domain.dx
        self.psi = psi.astype(np.float32)
        self.sqrt_kappa = math.sqrt(self.kappa)
        gx = 0.5 * (np.roll(psi, -1, 0) - np.roll(psi, 1, 0)) / self.hx
        gy = 0.5 * (np.roll(psi, -1, 1) - np.roll(psi, 1, 1)) / self.hy
        self.norm_grad_psi = (np.sqrt(gx**2 + gy**2) / psi).astype(np.float32)
        self.left_half = side.astype(np.float32)
        self._dev = {}

    def _on(self, device):
        import torch

        k = str(device)
        if k not in self._dev:
            self._dev[k] = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in (self.psi, self.norm_grad_psi, self.left_half))
        return self._dev[k]

    def _launch(self, y, mob, cos_a, cos_b, flux):
        import torch

        from .. import _lib

        psi, ngp, lh = self._on(y.device)
        with torch.no_grad():
            fval = self.f(y).to(torch.float32).contiguous()
            muval = self.mu(y).to(torch.float32).contiguous()
            mobv = mob(y).to(torch.float32).contiguous()
        d = _lib.SbmDesc(kind=self._sbm_kind, nx=y.shape[1], ny=y.shape[2], hx=float(self.hx), hy=float(self.hy), kappa=float(self.kappa))
        out = torch.empty_like(y)
        work = torch.empty_like(y) if self._sbm_kind == 0 else None
        vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()
        with _lib.device_of(y):
            _lib.check(_lib.load().pdeopt_sbm_rhs_batched(ctypes.byref(d), vp(y), vp(fval), vp(muval), vp(mobv), vp(psi), vp(ngp), vp(lh),
                                                          float(cos_a), float(cos_b), float(flux), vp(work), vp(out), y.shape[0],
                                                          _lib.stream_ptr(y)))
        return out


@dataclasses.dataclass
class CahnHilliard2DSmoothedBoundary(_SmoothedBoundary2D):
    """du/dt = (1/psi) div(psi D grad mu) + |grad psi|/psi J_n  (cahn_hilliard.py:203-289)."""

    domain: Domain
    kappa: float
    f: Any
    mu: Any
    D: Any
    theta: Any
    flux: Any
    derivs: str = "fd"
    _sbm_kind = 0

    def __post_init__(self):
        side = np.zeros(self.domain.points)
        side[:50, :] = 1.0  # cahn_hilliard.py:255-256
        self._setup(side)

    def rhs(self, state, t=0.0):
        single = state.dim() == 2
        y = (state.unsqueeze(0) if single else state).contiguous()
        th = float(self.theta(t))
        out = self._launch(y, self.D, math.cos(th), math.cos(math.pi - th), float(self.flux(t)))
        return out[0] if single else out

    rhs_fd = rhs


@dataclasses.dataclass
class AllenCahn2DSmoothedBoundary(_SmoothedBoundary2D):
    """du/dt = -R(u) mu with the boundary terms of the smoothed-boundary method (allen_cahn.py:87-159)."""

    domain: Domain
    kappa: float
    f: Any
    mu: Any
    R: Any
    theta: Any
    derivs: str = "fd"
    _sbm_kind = 1

    def __post_init__(self):
        side = np.zeros(self.domain.points)
        side[:, :100] = 1.0  # allen_cahn.py:136-137
        self._setup(side)

    def rhs(self, state, t=0.0):
        single = state.dim() == 2
        y = (state.unsqueeze(0) if single else state).contiguous()
        out = self._launch(y, self.R, math.cos(float(self.theta(t))), 0.0, 0.0)
        return out[0] if single else out

    rhs_fd = rhs
