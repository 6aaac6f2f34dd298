"""Advection-diffusion on a periodic 2-D grid (recovered equation).

The reference imports `AdvectionDiffusion2D` (notebooks/run_advection_diffusion.ipynb cell 0) and
quotes a deleted `AdvectionDiffusionEnv` (notebooks/test_pde_RL.ipynb:129), but the class is
absent from the tree.  Its semantics were recovered from the surviving fixture
notebooks/reference.npy:  du/dt = -div(v u) + D lap(u) with Fourier-spectral derivatives.
The class follows the reference's equation protocol (dataclass fields, class-level
`fft / ifft / fourier_symbol`, `rhs(state, t)`) so it pairs with SemiImplicitFourierSpectral;
the IMEX symbol sigma = D (2 pi)^2 |k|^2 (stiff linear part, A = 1) is our definition."""
import dataclasses
from typing import Any

import numpy as np

from ..domains import Domain
from ..functions import GaussianVelocity
from .base_eq import BaseEquation
from .phase_field import _symbols, spatial_fft, spatial_ifft


@dataclasses.dataclass
class AdvectionDiffusion2D(BaseEquation):
    domain: Domain
    velocity: Any  # GaussianVelocity (the enumerated family) — `advection(t, x, y)` in the notebook
    D: float
    fft = None
    ifft = None
    fourier_symbol = None

    def __post_init__(self):
        self.two_pi_i_kx, self.two_pi_i_ky, self.two_pi_i_k_2 = _symbols(self.domain)
        self.fft, self.ifft = spatial_fft(2), spatial_ifft(2)
        self.fourier_symbol = (-np.complex64(self.D) * self.two_pi_i_k_2).astype(np.complex64)
        self._tables = {}

    @property
    def fused(self):
        return isinstance(self.velocity, GaussianVelocity)

    def ad_desc(self):
        from .. import _lib

        d = _lib.AdDesc()
        d.nx, d.ny = self.domain.points
        d.lo_x, d.lo_y = self.domain.box[0][0], self.domain.box[1][0]
        d.hx, d.hy = self.domain.dx
        return d

    def spectral_tables(self, A=1.0):
        """Host table block of pdeopt_ad_rollout_* (include/pdeopt_b200.h)."""
        from ..fused import fold_symbol

        nx, ny = self.domain.points
        tab_a = fold_symbol(self.fourier_symbol, A).ravel()
        tab_l = fold_symbol(-(np.complex64(self.D) * self.two_pi_i_k_2), 1.0).ravel()
        kx = self.two_pi_i_kx[:, 0].imag.astype(np.float32).copy()
        ky = self.two_pi_i_ky[0, :].imag.astype(np.float32).copy()
        kx[nx // 2] = 0.0  # `.real` of the full complex transform kills the odd multipliers there
        ky[ny // 2] = 0.0
        return np.ascontiguousarray(np.concatenate([tab_a, tab_l, kx, ky]).astype(np.float32))

    def tables_on(self, device, A=1.0):
        import torch

        key = (str(device), float(A))
        if key not in self._tables:
            self._tables[key] = torch.from_numpy(self.spectral_tables(A)).to(device)
        return self._tables[key]

    def control_block(self, batch, device, nseg=1):
        return self.velocity.control_block(batch, nseg, device)

    def rhs(self, state, t=0.0):
        """f = -div(v u) + D lap(u) on CUDA float32 tensors ([nx,ny] or [B,nx,ny]), evaluated with the fused kernel as
        y1 - y0 of ONE explicit step of length 1 (A = 0 makes the IMEX filter the identity, so y1 = y0 + f exactly as the
        kernel forms it).  With dt = 1 the subtraction loses at most an ulp of max(|y0|, |f|) — a step of length 2^-10,
        as used before, amplified the rounding of y1 by 1/dt.  Not differentiable (no autograd node)."""
        import torch

        from ..adjoint import ad_rollout

        single = state.dim() == 2
        y = (state.unsqueeze(0) if single else state).contiguous()
        with torch.no_grad():
            ctrl = self.control_block(y.shape[0], y.device)
            y1 = ad_rollout(self, y.detach(), ctrl.detach(), np.asarray([0.0, 1.0], dtype=np.float32), A=0.0)
            f = y1 - y
        return f[0] if single else f
