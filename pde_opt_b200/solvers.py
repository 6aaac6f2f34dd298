"""Steppers.  Mirror of pde_opt/numerics/solvers.py: same class names, constructor fields,
`required_equation_attrs` protocol and `step(...)` signature / return tuple, with the
arithmetic running in the fused sm_100a kernels.

Differences a user can see (documented, not hidden):
* arrays are torch CUDA float32 tensors, optionally with a leading batch axis;
* `terms` is this module's ODETerm.  `ODETerm(eq)` (an equation object) takes the fused path;
  `ODETerm(callable)` evaluates the vector field first (solvers.py:59) and filters it in a
  second launch (the path for closures outside the enumerated families);
* `y_error` (solvers.py:61,65) is only materialised when the solver is built with
  `with_error=True` (only adaptive controllers consume it; the reference's drivers use
  ConstantStepSize)."""
import dataclasses
from typing import Any, Callable, Optional

import numpy as np
import torch

from .fused import SifsPlan, fold_symbol


class RESULTS:
    """Subset of diffrax.RESULTS used on this path."""
    successful = 0
    max_steps_reached = 1


class ODETerm:
    """Stand-in for diffrax.ODETerm: wraps f(t, y, args) or an equation object."""

    def __init__(self, vector_field):
        self.equation = vector_field if hasattr(vector_field, "rhs") and not callable(vector_field) else None
        self.vector_field = vector_field

    def vf(self, t, y, args=None):
        if self.equation is not None:
            return self.equation.rhs(y, t)
        return self.vector_field(t, y, args)


class LocalLinearInterpolation:
    """diffrax.LocalLinearInterpolation: y0 + (y1 - y0) * (t - t0) / (t1 - t0)."""

    def __init__(self, t0, t1, y0, y1):
        self.t0, self.t1, self.y0, self.y1 = t0, t1, y0, y1

    def evaluate(self, t):
        w = (np.float32(t) - np.float32(self.t0)) / (np.float32(self.t1) - np.float32(self.t0))
        return self.y0 + (self.y1 - self.y0) * float(w)


@dataclasses.dataclass
class SemiImplicitFourierSpectral:
    """y1 = y0 + dt * Re ifft( fft(f(y0)) / (1 + A dt symbol) )   (solvers.py:23-73)."""

    A: float
    fourier_symbol: Any
    fft: Optional[Callable] = None
    ifft: Optional[Callable] = None
    with_error: bool = False

    required_equation_attrs = ["fourier_symbol", "fft", "ifft"]  # solvers.py:42
    term_structure = ODETerm
    interpolation_cls = LocalLinearInterpolation

    def __post_init__(self):
        self._is3d = np.ndim(self.fourier_symbol) == 3
        # 2-D: folded quadrant table for the fused kernels; 3-D: position-ordered table, built lazily
        self._quad = None if self._is3d else fold_symbol(self.fourier_symbol, self.A)
        self._sym_dev = {}
        self._filter_plan = None

    def symbol_pos_on(self, device):
        """A * fourier_symbol as float32 [nx, ny, nz/2+1]: the line-FFT engine's position order along
        x and y, natural kz = 0..nz/2 along z (the z transform of the real field is real-to-half)."""
        key = ("pos", str(device))
        if key not in self._sym_dev:
            from .linefft import to_position_order

            s = np.asarray(self.fourier_symbol)
            if np.iscomplexobj(s):
                if np.abs(s.imag).max() > 1e-6 * max(1.0, np.abs(s.real).max()):
                    raise ValueError("fourier_symbol must be real for the semi-implicit path")
                s = s.real
            s = (np.float32(self.A) * s.astype(np.float32)).astype(np.float32)
            s = to_position_order(s, (0, 1))[:, :, : s.shape[2] // 2 + 1]
            self._sym_dev[key] = torch.from_numpy(np.ascontiguousarray(s)).to(device)
        return self._sym_dev[key]

    def _rollout3d(self, terms, dts, y0, out=None, times=None):
        eq = getattr(terms, "equation", None)
        single = y0.dim() == 3
        y = (y0.unsqueeze(0) if single else y0).contiguous()
        if eq is not None and getattr(eq, "_kind", None) == "ch3d" and eq.fused:
            y1 = eq.plan().step(y, dts, self.symbol_pos_on(y.device), out=out)
            return y1[0] if single else y1
        # unfused 3-D path (derivs="fourier", or any vector field): terms.vf evaluated by the caller's
        # code, then y1 = y0 + dt Re ifft(fft(f0) / (1 + A dt symbol)) on the line-FFT engine (solvers.py:59-63)
        from .linefft import fftn, ifftn

        key = ("nat", str(y.device))
        if key not in self._sym_dev:
            s_ = np.asarray(self.fourier_symbol)
            self._sym_dev[key] = torch.from_numpy((np.float32(self.A) * s_.real.astype(np.float32)).astype(np.float32)).to(y.device)
        sym = self._sym_dev[key]
        t = 0.0 if times is None else float(times[0])
        for k, dt in enumerate(dts):
            f0 = terms.vf(t if times is None else float(times[k]), y[0] if single else y, None)
            f0 = (f0.unsqueeze(0) if single else f0).contiguous()
            g = ifftn((fftn(f0, (1, 2, 3)) / (1.0 + float(dt) * sym)).contiguous(), (1, 2, 3)).real
            y = y + float(dt) * g
        if out is not None:
            out.copy_(y)
            y = out
        return y[0] if single else y

    def order(self, terms):
        return 1

    def init(self, terms, t0, t1, y0, args):
        return None

    def func(self, terms, t0, y0, args):
        return terms.vf(t0, y0, args)

    def symbol_on(self, device):
        key = str(device)
        if key not in self._sym_dev:
            self._sym_dev[key] = torch.from_numpy(self._quad).to(device)
        return self._sym_dev[key]

    def _rollout_ad(self, eq, times, y0, out=None):
        """Advection-diffusion (recovered equation): the fused forward kernel (differentiable)."""
        from .adjoint import ad_rollout

        single = y0.dim() == 2
        y = (y0.unsqueeze(0) if single else y0).contiguous()
        y1 = ad_rollout(eq, y, eq.control_block(y.shape[0], y.device), times, A=float(self.A))
        if out is not None:
            out.copy_(y1)
            y1 = out
        return y1[0] if single else y1

    def filter_plan(self, shape):
        """The plan of the unfused path (caller-evaluated vector field + pdeopt_sifs_filter_batched) for a grid."""
        if self._filter_plan is None:
            nx, ny = shape
            self._filter_plan = SifsPlan("ch2d", nx, ny, (0.0, 0.0), (1.0, 1.0), 0.0)
        return self._filter_plan

    def _plan_for(self, terms, shape):
        eq = getattr(terms, "equation", None)
        if eq is not None and getattr(eq, "fused", False):
            return eq.plan(), eq
        return self.filter_plan(shape), None

    def step(self, terms, t0, t1, y0, args=None, solver_state=None, made_jump=False):
        del solver_state, made_jump
        dt = np.float32(np.float32(t1) - np.float32(t0))  # solvers.py:58 in the working precision
        if type(getattr(terms, "equation", None)).__name__ == "AdvectionDiffusion2D":
            y1 = self._rollout_ad(terms.equation, np.asarray([t0, t1], dtype=np.float32), y0)
            y_error = (y1 - (y0 + float(dt) * terms.vf(t0, y0, args))) if self.with_error else None
            return y1, y_error, dict(y0=y0, y1=y1), None, RESULTS.successful
        if self._is3d:
            y1 = self._rollout3d(terms, [dt], y0)
            y_error = None
            if self.with_error:
                y_error = y1 - (y0 + float(dt) * terms.vf(t0, y0, args))  # solvers.py:61,65
            return y1, y_error, dict(y0=y0, y1=y1), None, RESULTS.successful
        single = y0.dim() == 2
        y = (y0.unsqueeze(0) if single else y0).contiguous()
        plan, eq = self._plan_for(terms, tuple(y.shape[-2:]))
        sym = self.symbol_on(y.device)
        f0 = None
        if eq is not None:
            y1 = plan.step(y, [dt], sym, ctrl=eq.control)
        else:
            f0 = terms.vf(t0, y0, args)
            f0 = (f0.unsqueeze(0) if single else f0).contiguous()
            y1 = plan.filter(y, f0, dt, sym)
        y_error = None
        if self.with_error:
            if f0 is None:
                f0 = plan.rhs(y, ctrl=eq.control)
            y_error = y1 - (y + float(dt) * f0)  # solvers.py:61,65
            y_error = y_error[0] if single else y_error
        y1 = y1[0] if single else y1
        dense_info = dict(y0=y0, y1=y1)
        return y1, y_error, dense_info, None, RESULTS.successful

    def rollout(self, terms, times, y0, ctrl=None, obs=None, obs_range=(0.0, 1.0), reward=None, out=None, nonfinite=None):
        """All steps between consecutive entries of `times` (host array, working precision) in as
        few launches as possible: the K-fused form of the diffeqsolve loop body."""
        times = np.asarray(times, dtype=np.float32)
        dts = (times[1:] - times[:-1]).astype(np.float32)
        if self._is3d:
            return self._rollout3d(terms, dts, y0, out=out, times=times)
        if type(getattr(terms, "equation", None)).__name__ == "AdvectionDiffusion2D":
            return self._rollout_ad(terms.equation, times, y0, out=out)
        single = y0.dim() == 2
        y = (y0.unsqueeze(0) if single else y0).contiguous()
        plan, eq = self._plan_for(terms, tuple(y.shape[-2:]))
        sym = self.symbol_on(y.device)
        if eq is not None:
            c = ctrl if ctrl is not None else eq.control
            y1 = plan.step(y, dts, sym, ctrl=c, obs=obs, obs_range=obs_range, reward=reward, out=out, nonfinite=nonfinite)
        else:
            y1 = y
            for k, dt in enumerate(dts):
                f0 = terms.vf(times[k], y1[0] if single else y1, None)
                f0 = (f0.unsqueeze(0) if single else f0).contiguous()
                y1 = plan.filter(y1, f0, dt, sym)
        return y1[0] if single else y1


def fold_complex_even(a):
    """(kx>=0, ky>=0) quadrant of a complex array that is even in each wavenumber, as float32
    [(nx/2+1)*(ny/2+1), 2]; None when the array is identically zero."""
    a = np.asarray(a)
    if not np.any(a):
        return None
    nx, ny = a.shape
    if not (np.allclose(a[1:, :], a[1:, :][::-1, :], rtol=1e-5) and np.allclose(a[:, 1:], a[:, 1:][:, ::-1], rtol=1e-5)):
        raise ValueError("A_term must be even in each wavenumber for the fused Strang path")
    q = a[: nx // 2 + 1, : ny // 2 + 1].astype(np.complex64)
    return np.ascontiguousarray(np.stack([q.real, q.imag], -1).astype(np.float32)).reshape(-1, 2)


@dataclasses.dataclass
class StrangSplitting:
    """Strang split-step method with global renormalisation (solvers.py:76-125)."""

    A_term: Any
    dx: float
    fft: Optional[Callable] = None
    ifft: Optional[Callable] = None
    time_scale: complex = 1.0

    required_equation_attrs = ["A_term", "dx", "fft", "ifft"]  # solvers.py:84
    term_structure = ODETerm
    interpolation_cls = LocalLinearInterpolation

    def __post_init__(self):
        a = np.asarray(self.A_term)
        self._fused128 = tuple(a.shape) == (128, 128)
        # 128x128: folded quadrant for the single-CTA kernel; other grids: full table for the line path
        self._a_host = fold_complex_even(a) if self._fused128 else None
        self._a_full = None
        if not self._fused128 and np.any(a):
            q = a.astype(np.complex64)
            self._a_full = np.ascontiguousarray(np.stack([q.real, q.imag], -1).astype(np.float32))
        self._a_dev = {}
        self._work = {}

    def order(self, terms):
        return 1

    def init(self, terms, t0, t1, y0, args):
        return None

    def func(self, terms, t0, y0, args):
        return NotImplementedError  # solvers.py:124-125 (sic)

    def _a_on(self, device):
        if self._a_host is None:
            return None
        key = str(device)
        if key not in self._a_dev:
            self._a_dev[key] = torch.from_numpy(self._a_host).to(device)
        return self._a_dev[key]

    def rollout(self, terms, times, y0, ctrl=None, out=None, **_):
        import ctypes

        from . import _lib

        eq = getattr(terms, "equation", None)
        if eq is None or not hasattr(eq, "gpe_desc"):
            raise NotImplementedError("StrangSplitting needs ODETerm(GPE2DTSControl equation)")
        if not getattr(eq, "fused", False):
            return self._rollout_callback_lights(eq, times, y0, out)
        times = np.asarray(times, dtype=np.float32)
        dts = np.ascontiguousarray((times[1:] - times[:-1]).astype(np.float32))
        single = y0.dim() == 3
        y = (y0.unsqueeze(0) if single else y0).contiguous()
        assert y.is_cuda and y.dtype == torch.float32 and y.shape[-1] == 2
        y1 = out if out is not None else torch.empty_like(y)
        a = self._a_on(y.device)
        c = ctrl if ctrl is not None else (eq.control_block(y.shape[0], y.device) if eq._light.amp != 0.0 else None)
        ts = complex(self.time_scale)
        desc = eq.gpe_desc()
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)
        if not self._fused128:
            # grids that do not fit one SM (256x256 complex64, BASELINE config 3): line-FFT path
            nx, ny = int(y.shape[1]), int(y.shape[2])
            key = (nx, ny, y.shape[0], str(y.device))
            if key not in self._work:
                n = int(lib.pdeopt_strang_lines_work_floats(nx, ny, y.shape[0]))
                self._work[key] = torch.empty(n, dtype=torch.float32, device=y.device)
            af = None
            if self._a_full is not None:
                if ("full", str(y.device)) not in self._a_dev:
                    self._a_dev[("full", str(y.device))] = torch.from_numpy(self._a_full).to(y.device)
                af = self._a_dev[("full", str(y.device))]
            st = lib.pdeopt_strang_lines_step_batched(
                ctypes.byref(desc), ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(y1.data_ptr()), y.shape[0], len(dts),
                dts.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(af.data_ptr()) if af is not None else None,
                float(ts.real), float(ts.imag), ctypes.c_void_p(c.data_ptr()) if c is not None else None,
                ctypes.c_void_p(self._work[key].data_ptr()), stream,
            )
            _lib.check(st)
            return y1[0] if single else y1
        done, src = 0, y
        while done < len(dts):
            k = min(_lib.MAX_FUSED_STEPS, len(dts) - done)
            st = lib.pdeopt_strang_step_batched(
                ctypes.byref(desc), ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(y1.data_ptr()), y.shape[0], k,
                dts[done:].ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(a.data_ptr()) if a is not None else None,
                float(ts.real), float(ts.imag), ctypes.c_void_p(c.data_ptr()) if c is not None else None, stream,
            )
            _lib.check(st)
            src = y1
            done += k
        return y1[0] if single else y1

    def _rollout_callback_lights(self, eq, times, y0, out=None):
        """Unfused path for a `lights(t, x, y)` callable outside the enumerated family (any time
        dependence): per numeric step the caller's callable is evaluated at the step's t0 — where the
        reference evaluates b = terms.vf(t0, y0), solvers.py:109 — and the field is handed to the Strang
        kernels of the line-FFT engine as an additive potential (pdeopt_strang_lines_step_batched_light)."""
        import ctypes

        from . import _lib

        times = np.asarray(times, dtype=np.float32)
        dts = np.ascontiguousarray((times[1:] - times[:-1]).astype(np.float32))
        single = y0.dim() == 3
        y = (y0.unsqueeze(0) if single else y0).contiguous()
        assert y.is_cuda and y.dtype == torch.float32 and y.shape[-1] == 2
        y1 = out if out is not None else torch.empty_like(y)
        nx, ny = int(y.shape[1]), int(y.shape[2])
        lib = _lib.load()
        key = ("cb", nx, ny, y.shape[0], str(y.device))
        if key not in self._work:
            n = int(lib.pdeopt_strang_lines_work_floats(nx, ny, y.shape[0]))
            self._work[key] = torch.empty(n, dtype=torch.float32, device=y.device)
        a = np.asarray(self.A_term)
        af = None
        if np.any(a):
            if ("full", str(y.device)) not in self._a_dev:
                q = a.astype(np.complex64)
                full = np.ascontiguousarray(np.stack([q.real, q.imag], -1).astype(np.float32))
                self._a_dev[("full", str(y.device))] = torch.from_numpy(full).to(y.device)
            af = self._a_dev[("full", str(y.device))]
        ts = complex(self.time_scale)
        desc = eq.gpe_desc()
        src = y
        with _lib.device_of(y):
            stream = _lib.stream_ptr(y)
            for k in range(len(dts)):
                light = eq.light_field(float(times[k]), y.device)
                st = lib.pdeopt_strang_lines_step_batched_light(
                    ctypes.byref(desc), ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(y1.data_ptr()), y.shape[0], 1,
                    dts[k:].ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(af.data_ptr()) if af is not None else None,
                    float(ts.real), float(ts.imag), None, ctypes.c_void_p(light.data_ptr()), 0,
                    ctypes.c_void_p(self._work[key].data_ptr()), stream,
                )
                _lib.check(st)
                src = y1
        return y1[0] if single else y1

    def step(self, terms, t0, t1, y0, args=None, solver_state=None, made_jump=False):
        del solver_state, made_jump
        y1 = self.rollout(terms, np.asarray([t0, t1], dtype=np.float32), y0)
        return y1, None, dict(y0=y0, y1=y1), None, RESULTS.successful


# ---- explicit solvers for equations without a linear symbol (the smoothed-boundary equations) ---------------------
# The reference passes diffrax solver classes straight through PDEModel (docs/notebooks/
# solving_pde_smoothed_boundary.ipynb: PDEModel(AllenCahn2DSmoothedBoundary, domain, dfx.Tsit5) with a PIDController).
# diffrax is not installable here; these are the same kind of object for this package's PDEModel: no required
# equation attributes, `step` returns the embedded error estimate.  The stages are evaluated by `terms.vf` (the
# equation's CUDA RHS); the stage combinations are torch axpy calls (host-orchestrated, off the semi-implicit hot
# path).  Tsit5's tableau is not reproduced from memory: Dopri5 (diffrax has it too) is the 5(4) pair provided.

class Euler:
    """Forward Euler (diffrax.Euler): order 1, no error estimate."""

    required_equation_attrs = []
    term_structure = ODETerm
    interpolation_cls = LocalLinearInterpolation

    def __init__(self):
        pass

    def order(self, terms):
        return 1

    def init(self, terms, t0, t1, y0, args):
        return None

    def func(self, terms, t0, y0, args):
        return terms.vf(t0, y0, args)

    def step(self, terms, t0, t1, y0, args=None, solver_state=None, made_jump=False):
        dt = float(np.float32(np.float32(t1) - np.float32(t0)))
        y1 = y0 + dt * terms.vf(t0, y0, args)
        return y1, None, dict(y0=y0, y1=y1), None, RESULTS.successful

    def rollout(self, terms, times, y0, out=None, **_):
        times = np.asarray(times, dtype=np.float32)
        y = y0
        for k in range(len(times) - 1):
            y = self.step(terms, times[k], times[k + 1], y)[0]
        if out is not None:
            out.copy_(y)
            return out
        return y


class Dopri5(Euler):
    """Dormand-Prince 5(4) (diffrax.Dopri5): seven stages, fifth-order solution, embedded fourth-order error estimate."""

    _C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
    _A = ((), (1 / 5,), (3 / 40, 9 / 40), (44 / 45, -56 / 15, 32 / 9), (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
          (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656), (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84))
    _B4 = (5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40)

    with_error = True

    def order(self, terms):
        return 5

    def step(self, terms, t0, t1, y0, args=None, solver_state=None, made_jump=False):
        t0f = float(np.float32(t0))
        dt = float(np.float32(np.float32(t1) - np.float32(t0)))
        k = []
        for s in range(7):
            ys = y0
            for a, ki in zip(self._A[s], k):
                if a != 0.0:
                    ys = ys + (dt * a) * ki
            k.append(terms.vf(t0f + self._C[s] * dt, ys, args))
        y1 = ys  # the seventh stage point is the fifth-order solution (first-same-as-last)
        b5 = self._A[6] + (0.0,)
        err = None
        for b, bs, ki in zip(b5, self._B4, k):
            if b != bs:
                term = (dt * (b - bs)) * ki
                err = term if err is None else err + term
        return y1, err, dict(y0=y0, y1=y1), None, RESULTS.successful
