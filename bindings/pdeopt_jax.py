"""jax.ffi registration + jax.custom_vjp wrappers over libpdeopt_jax_ffi.so (built from pdeopt_jax_ffi.cc).

The module a pde-opt maintainer drops next to `pde_opt/numerics/solvers.py`; it needs jax >= 0.5 (jax.ffi) and a
CUDA jaxlib.  Neither exists in the image this repository was developed in, so nothing here is imported by the
package or its tests except `tests/test_bindings_source.py`, which checks it against include/pdeopt_b200.h
textually.  The torch-side twin that IS tested is pde_opt_b200/adjoint.py (torch.autograd.Function in the role
of custom_vjp); the argument order and semantics below are the same.

Usage inside the reference (replaces `diffrax.diffeqsolve(...)` at pde_env.py:293-303 / pde_model.py:120-134):

    from pdeopt_jax import B200Stepper
    stepper = B200Stepper.for_equation(eq, A=0.5)          # eq: CahnHilliard2DPeriodic / AllenCahn2DPeriodic
    y1 = stepper.rollout(y0, dts)                           # [B, nx, ny] -> [B, nx, ny], K = len(dts) fused steps
"""
import ctypes
import os

import numpy as np

try:  # guarded: the repository's own image has no jax
    import jax
    import jax.numpy as jnp
except ImportError as exc:  # pragma: no cover
    raise ImportError("bindings/pdeopt_jax.py needs jax (jax.ffi); see bindings/pdeopt_jax_ffi.cc for the build line") from exc

_HERE = os.path.dirname(os.path.abspath(__file__))
_ABI = ctypes.CDLL(os.environ.get("PDEOPT_B200_LIB", os.path.join(_HERE, "..", "pde_opt_b200", "libpdeopt_b200.so")),
                   mode=ctypes.RTLD_GLOBAL)
_FFI = ctypes.CDLL(os.environ.get("PDEOPT_JAX_FFI_LIB", os.path.join(_HERE, "libpdeopt_jax_ffi.so")))

_TARGETS = {
    "pdeopt_sifs_step": "PdeoptSifsStep", "pdeopt_sifs_filter": "PdeoptSifsFilter", "pdeopt_rhs": "PdeoptRhs",
    "pdeopt_pf_adjoint": "PdeoptPfAdjoint", "pdeopt_pf_tangent": "PdeoptPfTangent", "pdeopt_rollout_fwd": "PdeoptRolloutFwd",
    "pdeopt_strang_step": "PdeoptStrangStep", "pdeopt_ad_fwd": "PdeoptAdFwd", "pdeopt_ad_bwd": "PdeoptAdBwd",
}
for _name, _sym in _TARGETS.items():
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_FFI, _sym)), platform="CUDA")

MAX_FUSED = 512  # PDEOPT_MAX_FUSED_STEPS


class _PlanDesc(ctypes.Structure):  # pdeopt_plan_desc
    _fields_ = [("kind", ctypes.c_int32), ("derivs", ctypes.c_int32), ("nx", ctypes.c_int32), ("ny", ctypes.c_int32),
                ("lo_x", ctypes.c_double), ("lo_y", ctypes.c_double), ("hx", ctypes.c_double), ("hy", ctypes.c_double),
                ("kappa", ctypes.c_double),
                ("mu_family", ctypes.c_int32), ("mu_ncoef", ctypes.c_int32), ("mu_coef", ctypes.c_double * 16),
                ("mob_family", ctypes.c_int32), ("mob_ncoef", ctypes.c_int32), ("mob_coef", ctypes.c_double * 16)]


def _check(status):
    if status != 0:
        _ABI.pdeopt_last_error.restype = ctypes.c_char_p
        raise RuntimeError(_ABI.pdeopt_last_error().decode())


def make_plan(kind, nx, ny, lo, h, kappa, mu_family, mu_coef, mob_family, mob_coef, derivs=0):
    """pdeopt_plan_create; returns the plan address as a Python int (the `plan` attribute of the FFI calls)."""
    d = _PlanDesc(kind=kind, derivs=derivs, nx=nx, ny=ny, lo_x=lo[0], lo_y=lo[1], hx=h[0], hy=h[1], kappa=kappa,
                  mu_family=mu_family, mu_ncoef=len(mu_coef), mob_family=mob_family, mob_ncoef=len(mob_coef))
    for i, c in enumerate(mu_coef):
        d.mu_coef[i] = float(c)
    for i, c in enumerate(mob_coef):
        d.mob_coef[i] = float(c)
    handle = ctypes.c_void_p()
    _check(_ABI.pdeopt_plan_create(ctypes.byref(d), ctypes.byref(handle)))
    return int(handle.value)


def fold_symbol(fourier_symbol, A):
    """[nx/2+1, ny/2+1] quadrant of A * Re(fourier_symbol) (the symbols of the SIFS-compatible equations are real
    and even in each wavenumber, cahn_hilliard.py:64-80)."""
    s = np.asarray(fourier_symbol).real
    return jnp.asarray((A * s[: s.shape[0] // 2 + 1, : s.shape[1] // 2 + 1]).astype(np.float32).ravel())


def sifs_steps(plan, y0, symbol, dts, ctrl=None, obs_range=(0.0, 1.0)):
    """K = len(dts) fused SemiImplicitFourierSpectral steps (solvers.py:56-70): returns (y1, obs_u8, reward)."""
    B = y0.shape[0]
    ctrl = jnp.zeros((B, 8), jnp.float32) if ctrl is None else ctrl
    out = (jax.ShapeDtypeStruct(y0.shape, jnp.float32), jax.ShapeDtypeStruct(y0.shape, jnp.uint8),
           jax.ShapeDtypeStruct((B, 2), jnp.float32))
    return jax.ffi.ffi_call("pdeopt_sifs_step", out)(y0, symbol, ctrl, plan=np.int64(plan), dts=np.asarray(dts, np.float32),
                                                      obs_lo=np.float32(obs_range[0]), obs_hi=np.float32(obs_range[1]))


def _pf_work_floats(plan, batch):
    _ABI.pdeopt_phasefield_adjoint_work_floats.restype = ctypes.c_int64
    return int(_ABI.pdeopt_phasefield_adjoint_work_floats(ctypes.c_void_p(plan), ctypes.c_int32(batch)))


def make_phasefield_rollout(plan_factory, symbol, dts):
    """Differentiable rollout y_K = Phi(y0, mu_coef, mob_coef) for the finite-difference Cahn-Hilliard / Allen-Cahn
    equations: jax.custom_vjp whose backward pass runs pdeopt_phasefield_adjoint_step in reverse over the saved
    states (what jax.grad(model.mse) differentiates, pde_model.py:274-323).  `plan_factory(mu_coef, mob_coef)`
    returns a plan address for concrete coefficients (the coefficients are the point of linearisation), so the
    rollout must be called with concrete (non-traced) coefficient values — as optimistix's BFGS loop does when
    the objective is not jitted."""
    dts = np.asarray(dts, np.float32)

    @jax.custom_vjp
    def rollout(y0, mu_coef, mob_coef):
        plan = plan_factory(np.asarray(mu_coef), np.asarray(mob_coef))
        y = y0
        for k0 in range(0, len(dts), MAX_FUSED):
            y = sifs_steps(plan, y, symbol, dts[k0:k0 + MAX_FUSED])[0]
        return y

    def fwd(y0, mu_coef, mob_coef):
        plan = plan_factory(np.asarray(mu_coef), np.asarray(mob_coef))
        out = (jax.ShapeDtypeStruct(y0.shape, jnp.float32), jax.ShapeDtypeStruct((len(dts),) + y0.shape, jnp.float32))
        y, traj = jax.ffi.ffi_call("pdeopt_rollout_fwd", out)(y0, symbol, plan=np.int64(plan), dts=dts, save_every=np.int64(1))
        return y, (list(traj), plan, mu_coef.shape, mob_coef.shape)  # the adjoint needs the state before every step

    def bwd(res, lam):
        states, plan, mu_shape, mob_shape = res
        B = lam.shape[0]
        work = jnp.zeros((_pf_work_floats(plan, B),), jnp.float32)
        gmu = jnp.zeros((B, 16), jnp.float64)
        gmob = jnp.zeros((B, 16), jnp.float64)
        out = (jax.ShapeDtypeStruct(lam.shape, jnp.float32), jax.ShapeDtypeStruct(gmu.shape, jnp.float64),
               jax.ShapeDtypeStruct(gmob.shape, jnp.float64))
        call = jax.ffi.ffi_call("pdeopt_pf_adjoint", out, input_output_aliases={4: 1, 5: 2})
        for u, dt in zip(reversed(states), dts[::-1]):
            lam, gmu, gmob = call(u, lam, symbol, work, gmu, gmob, plan=np.int64(plan), dt=np.float32(dt))
        return lam, gmu.sum(0)[: mu_shape[0]].astype(jnp.float32), gmob.sum(0)[: mob_shape[0]].astype(jnp.float32)

    rollout.defvjp(fwd, bwd)
    return rollout


def make_ad_rollout(geom, tables, dts, hold, batch, nx, ny):
    """Differentiable advection-diffusion rollout y_K = Phi(y0, ctrl), ctrl [B, nseg, 4] = (cx, cy, p0, p1):
    forward saves the per-step states (internal layout), backward = pdeopt_ad_rollout_bwd."""
    dts = np.asarray(dts, np.float32)
    geom = np.asarray(geom, np.float64)  # (lo_x, lo_y, hx, hy)
    stride = 2 * ((batch + 1) // 2) * nx * ny

    def _fwd_call(y0, ctrl, save):
        out = (jax.ShapeDtypeStruct(y0.shape, jnp.float32),
               jax.ShapeDtypeStruct((len(dts), stride) if save else (1,), jnp.float32))
        return jax.ffi.ffi_call("pdeopt_ad_fwd", out)(y0, tables, ctrl, dts=dts, geom=geom, hold=np.int64(hold), step0=np.int64(0))

    @jax.custom_vjp
    def rollout(y0, ctrl):
        return _fwd_call(y0, ctrl, False)[0]

    def fwd(y0, ctrl):
        y1, traj = _fwd_call(y0, ctrl, True)
        return y1, (traj, ctrl)

    def bwd(res, lam1):
        traj, ctrl = res
        out = (jax.ShapeDtypeStruct(lam1.shape, jnp.float32), jax.ShapeDtypeStruct(ctrl.shape, jnp.float32))
        lam0, gctrl = jax.ffi.ffi_call("pdeopt_ad_bwd", out, input_output_aliases={4: 1})(
            traj, lam1, tables, ctrl, jnp.zeros_like(ctrl), dts=dts, geom=geom, hold=np.int64(hold), step0=np.int64(0))
        return lam0, gctrl

    rollout.defvjp(fwd, bwd)
    return rollout


class B200Stepper:
    """What `PDEEnv.step` / `PDEModel.solve` hold instead of a diffrax solver + diffeqsolve call."""

    def __init__(self, plan, symbol):
        self.plan, self.symbol = plan, symbol

    @classmethod
    def for_equation(cls, eq, A, mu=(1, (3.0,)), mob=(1, ())):
        """eq: CahnHilliard2DPeriodic (kind 0) or AllenCahn2DPeriodic (kind 1) of the reference; mu / mob are
        (family id, coefficients) pairs of include/pdeopt_b200.h (the closures cannot cross a C ABI)."""
        kind = 0 if type(eq).__name__.startswith("CahnHilliard") else 1
        nx, ny = eq.domain.points
        lo = (eq.domain.box[0][0], eq.domain.box[1][0])
        plan = make_plan(kind, nx, ny, lo, eq.domain.dx, float(eq.kappa), mu[0], mu[1], mob[0], mob[1])
        return cls(plan, fold_symbol(eq.fourier_symbol, A))

    def rollout(self, y0, dts, ctrl=None):
        y = y0
        for k0 in range(0, len(dts), MAX_FUSED):
            y = sifs_steps(self.plan, y, self.symbol, dts[k0:k0 + MAX_FUSED], ctrl)[0]
        return y
