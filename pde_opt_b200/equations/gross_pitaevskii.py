"""Gross-Pitaevskii equation with time splitting and control.

Mirror of pde_opt/numerics/equations/gross_pitaevskii.py:18-81 (GPE2DTSControl): same dataclass
fields and class-level `fft / ifft / A_term / dx`; the state is [N, N, 2] (re, im) float32.
`A_term` is identically zero exactly as shipped (`0.5j * two_pi_i_k_2 * 0.0`, line 62);
the solver honours whatever array it is given."""
import ctypes
import dataclasses
from typing import Any, Callable

import numpy as np

from ..domains import Domain
from ..functions import GaussianLight
from .base_eq import TimeSplittingEquation
from .phase_field import _fft_marker

hbar = 1.05e-34  # J*s
mass_Na23 = 3.8175406e-26  # kg
a0 = 5.29177210903e-11  # Bohr radius


@dataclasses.dataclass
class GPE2DTSControl(TimeSplittingEquation):
    domain: Domain
    k: float
    e: float
    lights: Any
    trap_factor: float = 1.0
    fft = None
    ifft = None
    A_term = None
    dx = None

    def __post_init__(self):
        self.dx = self.domain.dx[0]  # :51
        kx, ky = self.domain.fft_mesh()
        self.two_pi_i_kx = (2j * np.pi * kx).astype(np.complex64)
        self.two_pi_i_ky = (2j * np.pi * ky).astype(np.complex64)
        self.two_pi_i_k_2 = self.two_pi_i_kx**2 + self.two_pi_i_ky**2
        self.fft, self.ifft = _fft_marker, _fft_marker
        self.xmesh, self.ymesh = self.domain.mesh()
        self.A_term = (np.complex64(0.5j) * self.two_pi_i_k_2 * np.complex64(0.0)).astype(np.complex64)  # :62
        self._light = self._recognize_lights(self.lights)

    def _recognize_lights(self, lights):
        """`lights(t, x, y)` is a user callable in the reference (:61).  The fused kernel supports
        no light (a callable returning zeros) and a Gaussian spot (functions.GaussianLight)."""
        if lights is None:
            return GaussianLight(0.0, 0.0, 0.0, 1.0)
        if isinstance(lights, GaussianLight):
            return lights
        try:
            v = np.asarray(lights(0.0, self.xmesh, self.ymesh), dtype=np.float64)
            if np.all(np.broadcast_to(v, self.xmesh.shape) == 0.0):
                return GaussianLight(0.0, 0.0, 0.0, 1.0)
        except Exception:
            pass
        return None

    @property
    def fused(self):
        return self._light is not None

    def control_block(self, batch, device):
        import torch

        c = torch.zeros((batch, 8), dtype=torch.float32, device=device)
        c[:, 1], c[:, 2], c[:, 3], c[:, 4] = self._light.amp, self._light.x0, self._light.y0, self._light.width
        return c

    def gpe_desc(self):
        from .. import _lib

        d = _lib.GpeDesc()
        d.nx, d.ny = self.domain.points
        d.lo_x, d.lo_y = self.domain.box[0][0], self.domain.box[1][0]
        d.hx, d.hy = self.domain.dx
        d.k, d.e, d.trap_factor = float(self.k), float(self.e), float(self.trap_factor)
        return d

    def A_terms(self, state, t):  # :64-65
        return self.A_term * 0.0

    def B_terms(self, state, t):
        """b = -i V(psi) as [N,N,2]; host/torch evaluation for inspection (the fused kernel
        evaluates it internally)."""
        import torch

        x = torch.as_tensor(self.xmesh, dtype=torch.float32, device=state.device)
        y = torch.as_tensor(self.ymesh, dtype=torch.float32, device=state.device)
        V = 0.5 * self.trap_factor * ((1 + self.e) * x**2 + (1 - self.e) * y**2)
        V = V + self._light(0.0, x, y) + self.k * (state[..., 0] ** 2 + state[..., 1] ** 2)
        return torch.stack([torch.zeros_like(V), -V], dim=-1)

    def rhs(self, state, t):  # :77-81
        return self.B_terms(state, t)
