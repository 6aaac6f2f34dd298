// Right-hand sides of the smoothed-boundary phase-field equations (arbitrary geometries through a smooth
// level-set psi): CahnHilliard2DSmoothedBoundary.rhs_fd (pde_opt/numerics/equations/cahn_hilliard.py:261-289) and
// AllenCahn2DSmoothedBoundary.rhs_fd (allen_cahn.py:142-159), stencils of utils/derivatives.py:24-66.
// The reference integrates them with explicit diffrax solvers (docs/notebooks/solving_pde_smoothed_boundary.ipynb:
// Tsit5 + PIDController), so this is an RHS kernel, not a semi-implicit step.  The pointwise closures f, mu, D / R are
// arbitrary callables in the reference and are evaluated by the caller on the whole batch (given fields).
//   inner = mu - (kappa / psi) div(psi_f grad_f u) - sqrt(kappa) |grad psi|/psi sqrt(2 f) (cos th lh + cos(pi - th) (1 - lh))
//   CH:  out = div(psi_f D_f grad_f inner) / psi + |grad psi|/psi flux(t)        (two passes)
//   AC:  out = -R inner   with the contact-angle term restricted to lh            (one pass)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdeopt {

struct SbmParams {
  int nx, ny, batch, eq;  // eq: 0 = Cahn-Hilliard, 1 = Allen-Cahn
  const float *u, *fval, *muval, *mob;  // [B][nx][ny]: state, f(u), mu(u), D(u) or R(u)
  const float *psi, *ngp, *lh;          // [nx][ny]: level set, |grad psi| / psi, side mask
  float* inner;                          // [B][nx][ny] scratch (CH)
  float* out;                            // [B][nx][ny]
  float inv_hx, inv_hy, kappa, sqrt_kappa, cos_a, cos_b, flux;
};

static __global__ void __launch_bounds__(256) sbm_pass1_kernel(const __grid_constant__ SbmParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny;
  if (i >= npts) return;
  const size_t o = (size_t)blockIdx.y * npts;
  const int r = i / p.ny, c = i - r * p.ny;
  const int rp = (r + 1 == p.nx) ? 0 : r + 1, rm = (r == 0) ? p.nx - 1 : r - 1;
  const int cp = (c + 1 == p.ny) ? 0 : c + 1, cm = (c == 0) ? p.ny - 1 : c - 1;
  const float* u = p.u + o;
  const float u0 = u[i], s0 = p.psi[i];
  // div( psi_face * grad_face(u) ): face value = average, face gradient = forward difference, divergence = backward difference
  const float fxp = 0.5f * (s0 + p.psi[rp * p.ny + c]) * ((u[rp * p.ny + c] - u0) * p.inv_hx);
  const float fxm = 0.5f * (p.psi[rm * p.ny + c] + s0) * ((u0 - u[rm * p.ny + c]) * p.inv_hx);
  const float fyp = 0.5f * (s0 + p.psi[r * p.ny + cp]) * ((u[r * p.ny + cp] - u0) * p.inv_hy);
  const float fym = 0.5f * (p.psi[r * p.ny + cm] + s0) * ((u0 - u[r * p.ny + cm]) * p.inv_hy);
  const float lap = (fxp - fxm) * p.inv_hx + (fyp - fym) * p.inv_hy;
  const float lh = p.lh[i];
  const float angle = (p.eq == 1) ? p.cos_a * lh : (p.cos_a * lh + p.cos_b * (1.0f - lh));
  const float inner = p.muval[o + i] - (p.kappa / s0) * lap - p.sqrt_kappa * p.ngp[i] * sqrtf(2.0f * p.fval[o + i]) * angle;
  if (p.eq == 1) p.out[o + i] = -p.mob[o + i] * inner;
  else p.inner[o + i] = inner;
}

static __global__ void __launch_bounds__(256) sbm_pass2_kernel(const __grid_constant__ SbmParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny;
  if (i >= npts) return;
  const size_t o = (size_t)blockIdx.y * npts;
  const int r = i / p.ny, c = i - r * p.ny;
  const int rp = (r + 1 == p.nx) ? 0 : r + 1, rm = (r == 0) ? p.nx - 1 : r - 1;
  const int cp = (c + 1 == p.ny) ? 0 : c + 1, cm = (c == 0) ? p.ny - 1 : c - 1;
  const float *in = p.inner + o, *D = p.mob + o;
  const float m0 = in[i], D0 = D[i], s0 = p.psi[i];
  auto face = [&](int j, float inv_h, bool plus) {
    const float sf = 0.5f * (s0 + p.psi[j]), Df = 0.5f * (D0 + D[j]);
    const float g = plus ? (in[j] - m0) * inv_h : (m0 - in[j]) * inv_h;
    return sf * Df * g;
  };
  const float div = (face(rp * p.ny + c, p.inv_hx, true) - face(rm * p.ny + c, p.inv_hx, false)) * p.inv_hx +
                    (face(r * p.ny + cp, p.inv_hy, true) - face(r * p.ny + cm, p.inv_hy, false)) * p.inv_hy;
  p.out[o + i] = div / s0 + p.ngp[i] * p.flux;
}

}  // namespace pdeopt
