"""Thin torch-facing wrapper over the C ABI (device pointers + stream in, status out).

torch is used for device memory and streams only; all arithmetic of the stepping path
happens inside libpdeopt_b200.so."""
import ctypes

import numpy as np
import torch

from . import _lib

MU_FAMILIES = {"double_well": 0, "log": 1, "legendre": 2, "legendre_logprior": 3}
MOB_FAMILIES = {"const": 0, "degenerate": 1, "one_plus_sq": 2, "legendre_exp": 3}


def fold_symbol(symbol, A=1.0):
    """Quadrant [nx/2+1, ny/2+1] of A * fourier_symbol (cahn_hilliard.py:74, solvers.py:62),
    after checking that the symbol is real and even in each wavenumber (so that the quadrant
    determines it).  float32, C-contiguous."""
    s = np.asarray(symbol)
    nx, ny = s.shape
    if np.iscomplexobj(s):
        if np.abs(s.imag).max() > 1e-6 * max(1.0, np.abs(s.real).max()):
            raise ValueError("fourier_symbol must be real for the fused SIFS path")
        s = s.real
    even = np.allclose(s[1:, :], s[1:, :][::-1, :], rtol=1e-5) and np.allclose(s[:, 1:], s[:, 1:][:, ::-1], rtol=1e-5)
    if not even:
        raise ValueError("fourier_symbol must be even in each wavenumber for the fused SIFS path")
    q = s[: nx // 2 + 1, : ny // 2 + 1].astype(np.float32)
    return np.ascontiguousarray(np.float32(A) * q)


def _ptr(t):
    if t is None:
        return ctypes.c_void_p()
    if isinstance(t, np.ndarray):
        return t.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(t.data_ptr())


class SifsPlan:
    """One fused-stepper plan: equation kind, grid, pointwise families (pdeopt_plan_create)."""

    def __init__(self, kind, nx, ny, lo, h, kappa, mu=("double_well", ()), mob=("const", (1.0,)), derivs="fd"):
        lib = _lib.load()
        d = _lib.PlanDesc()
        d.kind = {"ch2d": _lib.KIND_CH2D, "ac2d": _lib.KIND_AC2D}[kind]
        d.derivs = {"fd": _lib.DERIVS_FD, "fourier": _lib.DERIVS_FOURIER}[derivs]
        d.nx, d.ny = nx, ny
        d.lo_x, d.lo_y = float(lo[0]), float(lo[1])
        d.hx, d.hy = float(h[0]), float(h[1])
        d.kappa = float(kappa)
        d.mu_family = MU_FAMILIES[mu[0]]
        d.mu_ncoef = len(mu[1])
        for i, c in enumerate(mu[1]):
            d.mu_coef[i] = float(c)
        d.mob_family = MOB_FAMILIES[mob[0]]
        d.mob_ncoef = len(mob[1])
        for i, c in enumerate(mob[1]):
            d.mob_coef[i] = float(c)
        self._h = ctypes.c_void_p()
        _lib.check(lib.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(self._h)))
        self.nx, self.ny = nx, ny
        self.table_len = int(lib.pdeopt_table_len(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().pdeopt_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def step(self, y0, dts, symbol, ctrl=None, obs=None, obs_range=(0.0, 1.0), reward=None, out=None, nonfinite=None):
        """K = len(dts) fused steps on y0 [B, nx, ny] float32 CUDA (pdeopt_sifs_step_batched).
        `symbol` is the folded A*symbol table on the device.  Returns y1 (out or new).
        `nonfinite`: optional int32 [B] CUDA tensor that receives 1 for every environment whose
        y1 holds a NaN / Inf (pdeopt_plan_set_nonfinite_flags)."""
        lib = _lib.load()
        assert y0.is_cuda and y0.dtype == torch.float32 and y0.is_contiguous()
        B = y0.shape[0]
        assert tuple(y0.shape[1:]) == (self.nx, self.ny)
        y1 = out if out is not None else torch.empty_like(y0)
        dts = np.ascontiguousarray(np.asarray(dts, dtype=np.float32))
        K = len(dts)
        assert symbol.is_cuda and symbol.dtype == torch.float32 and symbol.numel() == self.table_len
        done, src = 0, y0
        with _lib.device_of(y0):
            stream = _lib.stream_ptr(y0)
            if nonfinite is not None:
                assert nonfinite.is_cuda and nonfinite.dtype == torch.int32 and nonfinite.numel() >= B
            _lib.check(lib.pdeopt_plan_set_nonfinite_flags(self._h, _ptr(nonfinite)))
            while done < K:
                k = min(_lib.MAX_FUSED_STEPS, K - done)
                last = done + k == K
                st = lib.pdeopt_sifs_step_batched(
                    self._h, _ptr(src), _ptr(y1), B, k, _ptr(dts[done:]), _ptr(symbol), _ptr(ctrl),
                    _ptr(obs) if last else None, float(obs_range[0]), float(obs_range[1]),
                    _ptr(reward) if last else None, stream,
                )
                _lib.check(st)
                src = y1
                done += k
        return y1

    def filter(self, y0, f0, dt, symbol, out=None):
        """One step with an externally evaluated vector field (pdeopt_sifs_filter_batched)."""
        lib = _lib.load()
        y1 = out if out is not None else torch.empty_like(y0)
        with _lib.device_of(y0):
            _lib.check(lib.pdeopt_plan_set_nonfinite_flags(self._h, None))
            _lib.check(lib.pdeopt_sifs_filter_batched(self._h, _ptr(y0), _ptr(f0), _ptr(y1), y0.shape[0], float(dt), _ptr(symbol),
                                                      _lib.stream_ptr(y0)))
        return y1

    def rollout_fwd(self, y0, dts, symbol, save_every=1, out=None):
        """len(dts) steps that also keep the state at the start of every `save_every`-th step
        (pdeopt_sifs_rollout_fwd): returns (y1, traj [ceil(K / save_every), B, nx, ny])."""
        lib = _lib.load()
        assert y0.is_cuda and y0.dtype == torch.float32 and y0.is_contiguous()
        dts = np.ascontiguousarray(np.asarray(dts, dtype=np.float32))
        K, B = len(dts), y0.shape[0]
        y1 = out if out is not None else torch.empty_like(y0)
        traj = torch.empty((-(-K // save_every),) + tuple(y0.shape), dtype=torch.float32, device=y0.device)
        with _lib.device_of(y0):
            _lib.check(lib.pdeopt_plan_set_nonfinite_flags(self._h, None))
            _lib.check(lib.pdeopt_sifs_rollout_fwd(self._h, _ptr(y0), _ptr(y1), B, K, _ptr(dts), _ptr(symbol), _ptr(traj),
                                                   int(save_every), _lib.stream_ptr(y0)))
        return y1, traj

    def rollout_bwd(self, traj, lam, dts, symbol, gmu, gmob):
        """Discrete adjoint of the len(dts) steps whose start states are traj [K, B, nx, ny], IN PLACE on the
        cotangent lam [B, nx, ny]; gmu / gmob [B, 16] float64 accumulate the coefficient cotangents
        (pdeopt_sifs_rollout_bwd: one fused launch per 512 steps on 128 x 128 grids)."""
        lib = _lib.load()
        dts = np.ascontiguousarray(np.asarray(dts, dtype=np.float32))
        B = lam.shape[0]
        assert traj.is_contiguous() and lam.is_contiguous() and traj.shape[0] == len(dts) and tuple(traj.shape[1:]) == tuple(lam.shape)
        work = None
        if (self.nx, self.ny) != (128, 128):
            key = (B, str(lam.device))
            if getattr(self, "_adj_work_key", None) != key:
                n = int(lib.pdeopt_phasefield_adjoint_work_floats(self._h, B))
                self._adj_work, self._adj_work_key = torch.empty(n, dtype=torch.float32, device=lam.device), key
            work = self._adj_work
        with _lib.device_of(lam):
            _lib.check(lib.pdeopt_sifs_rollout_bwd(self._h, _ptr(traj), _ptr(lam), _ptr(lam), B, len(dts), _ptr(dts), _ptr(symbol),
                                                   _ptr(work), _ptr(gmu), _ptr(gmob), _lib.stream_ptr(lam)))
        return lam

    def tangent_steps(self, traj, v, dts, dmu, dmob, symbol):
        """Forward-mode tangents v [ndir, B, nx, ny] advanced IN PLACE through the len(dts) steps whose start
        states are traj [K, B, nx, ny] (pdeopt_phasefield_tangent_steps); dmu / dmob [ndir, 16] float32."""
        lib = _lib.load()
        dts = np.ascontiguousarray(np.asarray(dts, dtype=np.float32))
        ndir, B = v.shape[0], v.shape[1]
        assert v.is_contiguous() and traj.is_contiguous() and traj.shape[0] == len(dts) and tuple(traj.shape[1:]) == tuple(v.shape[1:])
        assert dmu.dtype == torch.float32 and dmob.dtype == torch.float32 and tuple(dmu.shape) == (ndir, _lib.MAX_COEF) == tuple(dmob.shape)
        key = (B, ndir, str(v.device))
        if getattr(self, "_tan_work_key", None) != key:
            n = int(lib.pdeopt_phasefield_tangent_work_floats(self._h, B, ndir))
            self._tan_work, self._tan_work_key = torch.empty(n, dtype=torch.float32, device=v.device), key
        with _lib.device_of(v):
            _lib.check(lib.pdeopt_phasefield_tangent_steps(self._h, _ptr(traj), _ptr(v), B, ndir, len(dts), _ptr(dts), _ptr(dmu.contiguous()),
                                                           _ptr(dmob.contiguous()), _ptr(symbol), _ptr(self._tan_work), _lib.stream_ptr(v)))
        return v

    def rhs(self, y, ctrl=None, out=None):
        lib = _lib.load()
        f = out if out is not None else torch.empty_like(y)
        with _lib.device_of(y):
            _lib.check(lib.pdeopt_rhs_batched(self._h, _ptr(y), _ptr(f), y.shape[0], _ptr(ctrl), _lib.stream_ptr(y)))
        return f

    def step_host(self, y0, dts, symbol, ctrl=None, obs=None, obs_range=(0.0, 1.0), reward=None, out=None):
        """Same through pdeopt_sifs_step_batched_host: NumPy (ideally pinned) buffers in/out."""
        lib = _lib.load()
        y1 = out if out is not None else np.empty_like(y0)
        dts = np.ascontiguousarray(np.asarray(dts, dtype=np.float32))
        assert len(dts) <= _lib.MAX_FUSED_STEPS
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.pdeopt_plan_set_nonfinite_flags(self._h, None))
        st = lib.pdeopt_sifs_step_batched_host(
            self._h, _ptr(y0), _ptr(y1), y0.shape[0], len(dts), _ptr(dts), _ptr(symbol), _ptr(ctrl), _ptr(obs),
            float(obs_range[0]), float(obs_range[1]), _ptr(reward), stream,
        )
        _lib.check(st)
        return y1


class Ch3dPlan:
    """CahnHilliard3DPeriodic on the line-FFT engine (pdeopt_ch3d_rhs / pdeopt_ch3d_step)."""

    def __init__(self, points, h, kappa, mu=("double_well", ()), mob=("const", (1.0,))):
        d = _lib.Ch3dDesc()
        d.nx, d.ny, d.nz = (int(p) for p in points)
        d.hx, d.hy, d.hz = (float(v) for v in h)
        d.kappa = float(kappa)
        d.mu_family, d.mu_ncoef = MU_FAMILIES[mu[0]], len(mu[1])
        for i, c in enumerate(mu[1]):
            d.mu_coef[i] = float(c)
        d.mob_family, d.mob_ncoef = MOB_FAMILIES[mob[0]], len(mob[1])
        for i, c in enumerate(mob[1]):
            d.mob_coef[i] = float(c)
        self.desc = d
        self.points = tuple(int(p) for p in points)
        self._work = {}

    def _workbuf(self, batch, device):
        key = (batch, str(device))
        if key not in self._work:
            n = int(_lib.load().pdeopt_ch3d_work_floats(ctypes.byref(self.desc), batch))
            self._work[key] = torch.empty(n, dtype=torch.float32, device=device)
        return self._work[key]

    def rhs(self, u, halo_lo=None, halo_hi=None, out=None):
        """rhs_fd of u [B, nx, ny, nz] (cahn_hilliard.py:177-200)."""
        lib = _lib.load()
        assert u.is_cuda and u.dtype == torch.float32 and u.is_contiguous() and tuple(u.shape[1:]) == self.points
        f = out if out is not None else torch.empty_like(u)
        work = self._workbuf(u.shape[0], u.device)
        with _lib.device_of(u):
            _lib.check(lib.pdeopt_ch3d_rhs(ctypes.byref(self.desc), _ptr(u), _ptr(halo_lo), _ptr(halo_hi), _ptr(work), _ptr(f), u.shape[0],
                                           _lib.stream_ptr(u)))
        return f

    def step(self, y0, dts, symbol_pos, out=None):
        """len(dts) semi-implicit steps of y0 [B, nx, ny, nz]; symbol_pos: A*symbol [nx, ny, nz/2+1],
        position order along x and y, natural kz along z (SemiImplicitFourierSpectral.symbol_pos_on)."""
        lib = _lib.load()
        assert y0.is_cuda and y0.dtype == torch.float32 and y0.is_contiguous() and tuple(y0.shape[1:]) == self.points
        assert symbol_pos.is_cuda and symbol_pos.dtype == torch.float32
        assert tuple(symbol_pos.shape) == (self.points[0], self.points[1], self.points[2] // 2 + 1)
        y1 = out if out is not None else torch.empty_like(y0)
        dts = np.ascontiguousarray(np.asarray(dts, dtype=np.float32))
        work = self._workbuf(y0.shape[0], y0.device)
        with _lib.device_of(y0):
            _lib.check(lib.pdeopt_ch3d_step(ctypes.byref(self.desc), _ptr(y0), _ptr(y1), y0.shape[0], len(dts), _ptr(dts), _ptr(symbol_pos),
                                            _ptr(work), _lib.stream_ptr(y0)))
        return y1

    def adjoint_step(self, u, lam, dt, symbol_pos, gmu, gmob):
        """In place on `lam`: cotangent after one step from state u -> cotangent before it
        (pdeopt_ch3d_adjoint_step); gmu / gmob [B, 16] accumulate the coefficient cotangents."""
        lib = _lib.load()
        B = u.shape[0]
        key = ("adj", B, str(u.device))
        if key not in self._work:
            n = int(lib.pdeopt_ch3d_adjoint_work_floats(ctypes.byref(self.desc), B))
            self._work[key] = torch.empty(n, dtype=torch.float32, device=u.device)
        with _lib.device_of(u):
            _lib.check(lib.pdeopt_ch3d_adjoint_step(ctypes.byref(self.desc), _ptr(u), _ptr(lam), _ptr(lam), B, float(dt), _ptr(symbol_pos),
                                                    _ptr(self._work[key]), _ptr(gmu), _ptr(gmob), _lib.stream_ptr(u)))
