#!/usr/bin/env python
"""Run the TRUE reference (acoh64/pde-opt on jax + diffrax) on the BASELINE configurations and write golden
vectors into tests/golden/ref_*.npz — the artefact that would pin the oracle (and through it the CUDA path) at the
north-star tolerances (1e-5 after one step, 1e-3 after 1000 steps, 1e-4 on gradients).

It cannot run in the image this repository was developed in: jax, diffrax, equinox and optimistix are not
installed and there is no network.  Run it wherever the reference is importable:

    pip install jax diffrax equinox optimistix          # CPU wheels are enough
    PYTHONPATH=/path/to/pde-opt python tools/make_reference_fixtures.py [--out tests/golden] [--quick]

and commit the files it writes; tests/test_reference_fixtures.py picks them up (oracle on CPU, CUDA path under
-m gpu) and skips while they are absent.  Initial conditions come from numpy.random.default_rng(env_index)
(SURVEY 8d): they are stored in the fixture, so the consumer needs no jax PRNG.

Cases (float32, the dtype of the benchmark; `steps` numeric steps of size dt through PDEModel.solve, i.e.
diffrax.diffeqsolve with ConstantStepSize and SaveAt(ts), pde_model.py:120-134):
  c1_ac64      Allen-Cahn 64 x 64, mu = c^3 - c, R = 1, kappa = 0.002, dt = 5e-6, A = 1      1 / 16 / 1000 steps
  c2_ch128     Cahn-Hilliard 128 x 128, log potential (w = 3), D = c(1-c), dt = 1e-6, A = 0.5 1 / 16 / 1000 steps
  c2_ch128_ps  same equation from a phase-separated tanh-interface state                      1 / 16 steps
  c3_gpe       GPE 2-D Strang split-step, 128 x 128 (tests/test_solvers.py:107-205 parameters) 1 / 16 steps, real + imaginary time
  c5_ch3d      Cahn-Hilliard 3-D 32^3, log potential, D = 0.15                                1 / 16 steps
  grad_ch64    d mse / d (Legendre coefficients of mu, D) on 64 x 64, 50 steps (jax.grad, RecursiveCheckpointAdjoint)
"""
import argparse
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    ap.add_argument("--quick", action="store_true", help="skip the 1000-step cases")
    args = ap.parse_args()
    try:
        import diffrax as dfx
        import jax
        import jax.numpy as jnp
    except ImportError as exc:
        sys.exit(f"make_reference_fixtures: the reference's dependencies are missing ({exc}); see the module docstring")
    jax.config.update("jax_platforms", "cpu")
    from pde_opt import PDEModel
    from pde_opt.numerics.domains import Domain
    from pde_opt.numerics.equations import AllenCahn2DPeriodic, CahnHilliard2DPeriodic, CahnHilliard3DPeriodic, GPE2DTSControl
    from pde_opt.numerics.functions import ChemicalPotentialLegendrePolynomials, DiffusionLegendrePolynomials
    from pde_opt.numerics.solvers import SemiImplicitFourierSpectral, StrangSplitting

    os.makedirs(args.out, exist_ok=True)
    H, KAPPA = 0.01, 0.002

    def dom(*n):
        return Domain(tuple(n), tuple((-k * H / 2, k * H / 2) for k in n), "dimensionless")

    def noise_ic(shape, env, mean=0.5, amp=0.01, lo=0.0, hi=1.0):
        return np.clip(mean + amp * np.random.default_rng(env).normal(size=shape), lo, hi).astype(np.float32)

    def save(name, **arrays):
        path = os.path.join(args.out, f"ref_{name}.npz")
        np.savez_compressed(path, **arrays)
        print("wrote", path, {k: np.asarray(v).shape for k, v in arrays.items()})

    def run(model, params, y0, dt, steps_list, solver_params):
        out = {}
        for k in steps_list:
            ts = jnp.asarray([0.0, k * dt], dtype=jnp.float32)
            ys = model.solve(params, jnp.asarray(y0), ts, solver_params, dt0=dt, max_steps=10 * k + 10)
            out[f"y_{k}"] = np.asarray(ys[-1])
        return out

    long = [] if args.quick else [1000]

    # ---- C1 Allen-Cahn 64^2 (notebooks/test_implicit.ipynb) ----
    d = dom(64, 64)
    y0 = (0.01 * np.random.default_rng(0).normal(size=(64, 64))).astype(np.float32)
    m = PDEModel(AllenCahn2DPeriodic, d, SemiImplicitFourierSpectral)
    p = {"kappa": KAPPA, "mu": lambda c: c**3 - c, "R": lambda c: jnp.ones_like(c), "derivs": "fd"}
    save("c1_ac64", y0=y0, dt=np.float32(5e-6), A=np.float32(1.0), **run(m, p, y0, 5e-6, [1, 16] + long, {"A": 1.0}))

    # ---- C2 Cahn-Hilliard 128^2 (notebooks/optimize_nn_script.py:15-41) ----
    d = dom(128, 128)
    m = PDEModel(CahnHilliard2DPeriodic, d, SemiImplicitFourierSpectral)
    p = {"kappa": KAPPA, "mu": lambda c: jnp.log(c / (1.0 - c)) + 3.0 * (1.0 - 2.0 * c), "D": lambda c: (1.0 - c) * c, "derivs": "fd"}
    y0 = noise_ic((128, 128), 0)
    save("c2_ch128", y0=y0, dt=np.float32(1e-6), A=np.float32(0.5), **run(m, p, y0, 1e-6, [1, 16] + long, {"A": 0.5}))
    x = np.asarray(d.axes()[0])
    ps = (0.5 + 0.45 * np.tanh((0.3 - np.abs(x)) / 0.02))[:, None] * np.ones((1, 128))
    ps = np.clip(ps + 0.005 * np.random.default_rng(1).normal(size=ps.shape), 0.03, 0.97).astype(np.float32)
    save("c2_ch128_ps", y0=ps, dt=np.float32(1e-6), A=np.float32(0.5), **run(m, p, ps, 1e-6, [1, 16], {"A": 0.5}))

    # ---- C3 GPE Strang (tests/test_solvers.py:107-205) ----
    n = 128
    d = Domain((n, n), ((-7.5, 7.5), (-7.5, 7.5)), "dimensionless")
    m = PDEModel(GPE2DTSControl, d, StrangSplitting)
    rng = np.random.default_rng(2)
    X, Y = np.meshgrid(np.asarray(d.axes()[0]), np.asarray(d.axes()[1]), indexing="ij")
    psi = np.exp(-(X**2 + Y**2) / 8.0) * (1.0 + 0.05 * rng.normal(size=X.shape)) * np.exp(0.3j * X)
    psi = psi / np.sqrt(np.sum(np.abs(psi) ** 2) * float(d.dx[0]) ** 2)
    y0 = np.stack([psi.real, psi.imag], -1).astype(np.float32)
    p = {"k": 3371.7, "e": 0.0, "lights": lambda t, x, y: jnp.zeros_like(x), "trap_factor": 1.0}
    for name, ts_ in (("imag", -1j), ("real", 1.0)):
        save(f"c3_gpe128_{name}", y0=y0, dt=np.float32(2e-4 * np.pi), time_scale=np.complex64(ts_),
             **run(m, p, y0, 2e-4 * np.pi, [1, 16], {"time_scale": ts_}))

    # ---- C5 Cahn-Hilliard 3-D 32^3 (docs/notebooks/optimization_3D.ipynb) ----
    d = dom(32, 32, 32)
    m = PDEModel(CahnHilliard3DPeriodic, d, SemiImplicitFourierSpectral)
    p = {"kappa": KAPPA, "mu": lambda c: jnp.log(c / (1.0 - c)) + 3.0 * (1.0 - 2.0 * c), "D": lambda c: 0.15 * jnp.ones_like(c), "derivs": "fd"}
    y0 = noise_ic((32, 32, 32), 3)
    save("c5_ch3d32", y0=y0, dt=np.float32(1e-6), A=np.float32(0.5), **run(m, p, y0, 1e-6, [1, 16], {"A": 0.5}))

    # ---- gradients: d mse / d Legendre coefficients (pde_model.py:274-323 through jax.grad) ----
    d = dom(64, 64)
    m = PDEModel(CahnHilliard2DPeriodic, d, SemiImplicitFourierSpectral)
    mu_c = jnp.asarray([0.1, -2.2, 0.3, 0.25], dtype=jnp.float32)
    d_c = jnp.asarray([-0.3, 0.2, -0.1], dtype=jnp.float32)
    y0s = np.stack([np.clip(0.5 + 0.1 * np.random.default_rng(10 + b).normal(size=(64, 64)), 0.05, 0.95) for b in range(2)]).astype(np.float32)
    ts = jnp.asarray([0.0, 25e-6, 50e-6], dtype=jnp.float32)
    target = jnp.asarray(np.stack([np.stack([y0s[b]] * 2) for b in range(2)]))  # [B, T-1, 64, 64]

    def loss(mu_c_, d_c_):
        params = {"kappa": KAPPA, "mu": ChemicalPotentialLegendrePolynomials(mu_c_, lambda x: jnp.log(x / (1.0 - x))),
                  "D": DiffusionLegendrePolynomials(d_c_), "derivs": "fd"}
        return m.mse(params, (jnp.asarray(y0s), target), {"A": 0.5}, ts, {}, 0.0, adjoint=dfx.RecursiveCheckpointAdjoint())

    val, (g_mu, g_d) = jax.value_and_grad(loss, argnums=(0, 1))(mu_c, d_c)
    save("grad_ch64", y0s=y0s, ts=np.asarray(ts), mu_coef=np.asarray(mu_c), d_coef=np.asarray(d_c), target=np.asarray(target),
         loss=np.float64(val), g_mu=np.asarray(g_mu), g_d=np.asarray(g_d), dt=np.float32(1e-6), A=np.float32(0.5))


if __name__ == "__main__":
    main()
