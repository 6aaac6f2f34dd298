"""Cahn-Hilliard and Allen-Cahn equations on periodic 2-D grids.

Mirror of pde_opt/numerics/equations/cahn_hilliard.py:30-109 and allen_cahn.py:26-84: same
dataclass fields, same class-level `fft / ifft / fourier_symbol` attributes (the solver
compatibility check is a class-level hasattr, pde_opt/utils.py:23-25), same symbols.
`rhs(state, t)` runs the fused-kernel RHS (pdeopt_rhs_batched) on CUDA tensors."""
import dataclasses
from typing import Any, Callable, Optional

import numpy as np

from ..domains import Domain
from ..functions import Closure, recognize
from .base_eq import BaseEquation


class _SpatialFFT:
    """`eq.fft` / `eq.ifft` of the reference are jnp.fft.fftn / ifftn (cahn_hilliard.py:72-73, allen_cahn.py:64-65,
    gross_pitaevskii.py:58-59): transforms over the spatial axes, natural (fftfreq) order, complex64 out.  Here
    the same callables on CUDA tensors, over the LAST `ndim` axes (leading axes are batch), running on the
    line-FFT engine (linefft.fftn / ifftn); the fused steppers do their transforms in-kernel and only use these
    attributes as the compatibility marker the reference's check looks for (utils.py:23-25)."""

    def __init__(self, ndim, inverse):
        self.ndim, self.inverse = int(ndim), bool(inverse)

    def __call__(self, x):
        from ..linefft import fftn, ifftn

        dims = tuple(range(x.dim() - self.ndim, x.dim()))
        return (ifftn if self.inverse else fftn)(x.contiguous(), dims)

    def __repr__(self):
        return f"<{'ifftn' if self.inverse else 'fftn'} over the last {self.ndim} axes (line-FFT engine)>"


def spatial_fft(ndim):
    return _SpatialFFT(ndim, False)


def spatial_ifft(ndim):
    return _SpatialFFT(ndim, True)


def _symbols(domain):
    kx, ky = domain.fft_mesh()
    two_pi_i_kx = (2j * np.pi * kx).astype(np.complex64)
    two_pi_i_ky = (2j * np.pi * ky).astype(np.complex64)
    k2 = two_pi_i_kx**2 + two_pi_i_ky**2
    return two_pi_i_kx, two_pi_i_ky, k2


def _on(device, a, cache, key):
    import torch

    k = (key, str(device))
    if k not in cache:
        cache[k] = torch.as_tensor(a, device=device)
    return cache[k]


def spectral_rhs_ch(eq, state, qs, k2):
    """CahnHilliard{2,3}DPeriodic.rhs_fourier (cahn_hilliard.py:82-87, :165-175) for grids without a
    fused kernel: the same expression on the line-FFT engine (linefft.fftn / ifftn) with the
    pointwise closures evaluated by torch.  `state` has a leading batch axis."""
    from ..linefft import fftn, ifftn

    dims = tuple(range(1, state.dim()))
    sh = fftn(state, dims)
    tmp = fftn(eq.mu(state).contiguous(), dims) - eq.kappa * k2 * sh
    Du = eq.D(state)
    acc = None
    for q in qs:
        term = q * fftn((Du * ifftn(q * tmp, dims)).contiguous(), dims)
        acc = term if acc is None else acc + term
    return ifftn(acc.contiguous(), dims).real.contiguous()


def spectral_rhs_ac(eq, state, k2):
    """AllenCahn2DPeriodic.rhs_fourier (allen_cahn.py:74-79) on the line-FFT engine."""
    from ..linefft import fftn, ifftn

    dims = tuple(range(1, state.dim()))
    mu = ifftn((fftn(eq.mu(state).contiguous(), dims) - eq.kappa * k2 * fftn(state, dims)).contiguous(), dims)
    return (-eq.R(state) * mu.real).contiguous()


class _PhaseField2D(BaseEquation):
    _kind = None

    def _setup(self, mob_name):
        if self.derivs not in ("fd", "fourier"):
            raise ValueError(f"Invalid derivative type: {self.derivs}")
        self.two_pi_i_kx, self.two_pi_i_ky, self.two_pi_i_k_2 = _symbols(self.domain)
        self.fft, self.ifft = spatial_fft(2), spatial_ifft(2)
        self._mu_c = recognize(self.mu, "mu")
        self._mob_c = recognize(getattr(self, mob_name), "mob")
        self._plan = None
        self.control = None  # optional [B, 8] control block (see include/pdeopt_b200.h)

    @property
    def fused(self):
        """True when mu and the mobility are enumerated families and a fused kernel exists for the
        derivative type (finite differences: every supported grid; 'fourier': 128x128)."""
        if self._mu_c is None or self._mob_c is None:
            return False
        return self.derivs == "fd" or tuple(self.domain.points) == (128, 128)

    def plan(self):
        if self._plan is None:
            if not self.fused:
                raise NotImplementedError(
                    "no fused kernel for this equation (non-enumerated mu/D closure or derivs='fourier'); "
                    "use the unfused solver path"
                )
            from ..fused import SifsPlan

            nx, ny = self.domain.points
            self._plan = SifsPlan(
                self._kind, nx, ny, (self.domain.box[0][0], self.domain.box[1][0]), self.domain.dx, self.kappa,
                self._mu_c.descriptor(), self._mob_c.descriptor(), self.derivs,
            )
        return self._plan

    def _rhs_given_mu(self, y):
        """rhs_fd for closures outside the enumerated families (PeriodicCNN / Mixer2d of functions_nn, or any callable
        on torch tensors): mu_h (and the mobility, unless it is an enumerated family) evaluated by the closure on the
        whole batch, the stencils by pdeopt_rhs_given_mu_batched (cahn_hilliard.py:89-109, allen_cahn.py:81-84)."""
        import ctypes

        import torch

        from .. import _lib
        from ..fused import SifsPlan

        mob_fn = self.D if self._kind == "ch2d" else self.R
        if getattr(self, "_gm_plan", None) is None:
            nx, ny = self.domain.points
            mob = self._mob_c.descriptor() if self._mob_c is not None else ("const", (1.0,))
            self._gm_plan = SifsPlan(self._kind, nx, ny, (self.domain.box[0][0], self.domain.box[1][0]), self.domain.dx, self.kappa,
                                     ("double_well", ()), mob, "fd")
        with torch.no_grad():
            muh = self.mu(y).to(torch.float32).contiguous()
            mob_v = None if self._mob_c is not None else mob_fn(y).to(torch.float32).contiguous()
        assert muh.shape == y.shape, "mu must map [B, nx, ny] to [B, nx, ny]"
        f = torch.empty_like(y)
        work = torch.empty((2,) + tuple(y.shape), dtype=torch.float32, device=y.device)
        vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()
        with _lib.device_of(y):
            _lib.check(_lib.load().pdeopt_rhs_given_mu_batched(self._gm_plan._h, vp(y), vp(muh), vp(mob_v), vp(f), y.shape[0], vp(work),
                                                               _lib.stream_ptr(y)))
        return f

    def rhs(self, state, t=0.0):
        """eq.rhs(state, t) on CUDA float32 tensors, [nx,ny] or [B,nx,ny]."""
        import ctypes

        import torch

        from .. import _lib

        single = state.dim() == 2
        y = state.unsqueeze(0) if single else state
        y = y.contiguous()
        if self.derivs == "fourier" and not self.fused:
            # no fused kernel for this grid / closure: the reference expression on the line-FFT engine
            c = self.__dict__.setdefault("_dev_cache", {})
            k2 = _on(y.device, self.two_pi_i_k_2, c, "k2")
            if self._kind == "ac2d":
                f = spectral_rhs_ac(self, y, k2)
            else:
                f = spectral_rhs_ch(self, y, [_on(y.device, self.two_pi_i_kx, c, "qx"), _on(y.device, self.two_pi_i_ky, c, "qy")], k2)
            return f[0] if single else f
        if self.derivs == "fd" and not self.fused:
            f = self._rhs_given_mu(y)
            return f[0] if single else f
        out = torch.empty_like(y)
        plan = self.plan()
        st = _lib.load().pdeopt_rhs_batched(
            plan._h, ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(out.data_ptr()), y.shape[0],
            ctypes.c_void_p(self.control.data_ptr()) if self.control is not None else ctypes.c_void_p(),
            ctypes.c_void_p(torch.cuda.current_stream(y.device).cuda_stream),
        )
        _lib.check(st)
        return out[0] if single else out


@dataclasses.dataclass
class CahnHilliard2DPeriodic(_PhaseField2D):
    """du/dt = div(D(u) grad(mu)), mu = mu_h(u) - kappa lap(u)  (cahn_hilliard.py:30-109)."""

    domain: Domain
    kappa: float
    mu: Any
    D: Any
    derivs: str = "fd"
    fft = None
    ifft = None
    fourier_symbol = None
    _kind = "ch2d"

    def __post_init__(self):
        self._setup("D")
        self.two_pi_i_k_4 = self.two_pi_i_k_2**2
        self.fourier_symbol = (np.complex64(self.kappa) * self.two_pi_i_k_4).astype(np.complex64)  # :74


@dataclasses.dataclass
class AllenCahn2DPeriodic(_PhaseField2D):
    """du/dt = -R(u) mu  (allen_cahn.py:26-84).

    The reference class defines no `fourier_symbol` (so it cannot be paired with
    SemiImplicitFourierSpectral there); BASELINE config 1 needs one, so we define
    sigma = -kappa (2 pi i k)^2 = kappa (2 pi)^2 |k|^2, the stiff linear part for R == 1."""

    domain: Domain
    kappa: float
    mu: Any
    R: Any
    derivs: str = "fd"
    fft = None
    ifft = None
    fourier_symbol = None
    _kind = "ac2d"

    def __post_init__(self):
        self._setup("R")
        self.fourier_symbol = (-np.complex64(self.kappa) * self.two_pi_i_k_2).astype(np.complex64)


@dataclasses.dataclass
class CahnHilliard3DPeriodic(BaseEquation):
    """3-D Cahn-Hilliard on a periodic box (cahn_hilliard.py:112-200): same fields and class-level
    `fft / ifft / fourier_symbol` as the reference.  Runs on the line-FFT engine (the field does not
    fit one SM).  `fourier_symbol` is built on first use (512^3 complex64 is 1 GB on the host)."""

    domain: Domain
    kappa: float
    mu: Any
    D: Any
    derivs: str = "fd"
    fft = None
    ifft = None
    _kind = "ch3d"

    def __post_init__(self):
        if self.derivs not in ("fd", "fourier"):
            raise ValueError(f"Invalid derivative type: {self.derivs}")
        self.fft, self.ifft = spatial_fft(3), spatial_ifft(3)
        self._mu_c = recognize(self.mu, "mu")
        self._mob_c = recognize(self.D, "mob")
        self._plan = None
        self._symbol = None

    @property
    def fourier_symbol(self):
        if self._symbol is None:
            kx, ky, kz = self.domain.fft_mesh()
            t = [(2j * np.pi * k).astype(np.complex64) for k in (kx, ky, kz)]
            k2 = t[0] ** 2 + t[1] ** 2 + t[2] ** 2  # cahn_hilliard.py:150-154
            self._symbol = (np.complex64(self.kappa) * k2**2).astype(np.complex64)  # :158
        return self._symbol

    @property
    def fused(self):
        return self._mu_c is not None and self._mob_c is not None and self.derivs == "fd"

    def plan(self):
        if self._plan is None:
            if not self.fused:
                raise NotImplementedError("no kernel for this 3-D equation (non-enumerated mu/D closure or derivs='fourier')")
            from ..fused import Ch3dPlan

            self._plan = Ch3dPlan(self.domain.points, self.domain.dx, self.kappa, self._mu_c.descriptor(), self._mob_c.descriptor())
        return self._plan

    def rhs(self, state, t=0.0):
        single = state.dim() == 3
        y = (state.unsqueeze(0) if single else state).contiguous()
        if self.derivs == "fourier":
            # cahn_hilliard.py:165-175 on the line-FFT engine (7 complex 3-D transforms)
            c = self.__dict__.setdefault("_dev_cache", {})
            if "q" not in c:
                kx, ky, kz = self.domain.fft_mesh()
                c["q"] = [(2j * np.pi * k).astype(np.complex64) for k in (kx, ky, kz)]
                c["k2n"] = c["q"][0] ** 2 + c["q"][1] ** 2 + c["q"][2] ** 2
            qs = [_on(y.device, q, c, f"q{i}") for i, q in enumerate(c["q"])]
            f = spectral_rhs_ch(self, y, qs, _on(y.device, c["k2n"], c, "k2"))
        else:
            f = self.plan().rhs(y)
        return f[0] if single else f
