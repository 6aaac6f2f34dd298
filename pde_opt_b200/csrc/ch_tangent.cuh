// Forward-mode tangent (JVP) of one semi-implicit step of the finite-difference phase-field equations
// (Cahn-Hilliard / Allen-Cahn 2-D) with respect to the state and the closure coefficients.
//
// This is the derivative PDEModel.train(method="least_squares") asks for: the reference runs optimistix's
// Levenberg-Marquardt over PDEModel.residuals with diffrax's ForwardMode adjoint (pde_model.py:404-428), i.e. it
// pushes one tangent per optimised coefficient through every SemiImplicitFourierSpectral.step (solvers.py:56-70).
//
// Forward step:  y1 = u + dt G f(u; theta),  G = F^-1 [1/(1 + A dt sigma)] F,
//   CH: f = div( D_face grad_face(mu) ),  AC: f = -R(u) mu,   mu = mu_h(u; theta_mu) - kappa lap(u)
//   (cahn_hilliard.py:89-109, allen_cahn.py:81-84, derivatives.py:8-66).
// Tangent for a direction (v, dtheta_mu, dtheta_D):
//   mu~ = mu_h'(u) v + sum_n dtheta_mu[n] d mu_h / d theta_n (u) - kappa lap(v)
//   D~  = D'(u) v + sum_n dtheta_D[n] d D / d theta_n (u)
//   CH: f~ = div( D~_face grad_face(mu) + D_face grad_face(mu~) ),   AC: f~ = -D~ mu - D mu~
//   v1 = v + dt G f~      (the filter G does not depend on theta; pdeopt_sifs_filter_batched applies it).
// Two streaming kernels per step over (direction, environment); the state-dependent fields mu and D are
// evaluated once per environment and shared by the directions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ch_adjoint.cuh"  // mu_h_prime, mob_prime, PDEOPT_ADJ_NCOEF

namespace pdeopt {

struct ChTanParams {
  int nx, ny, batch, ndir, eq;  // eq: 0 = Cahn-Hilliard, 1 = Allen-Cahn
  const float* u;     // [B][nx][ny] state at the start of the step
  const float* v;     // [ndir][B][nx][ny] tangent of the state
  const float* dmu;   // [ndir][16] direction in mu_coef
  const float* dmob;  // [ndir][16] direction in mob_coef
  float* mu;          // [B][nx][ny] scratch
  float* dd;          // [B][nx][ny] scratch: D(u)
  float* mut;         // [ndir][B][nx][ny] scratch: mu~
  float* ddt;         // [ndir][B][nx][ny] scratch: D~
  float* ft;          // [ndir][B][nx][ny] out: f~
  float inv_hx, inv_hy, inv_hx2, inv_hy2, kappa;
  PointwiseParams pw;
};

// sum_n d[n] * d mu_h / d theta_n (c): the families' coefficient derivatives (see ch_adj_out_kernel)
__device__ __forceinline__ float mu_h_dtheta(float c, const PointwiseParams& pw, const float* __restrict__ d) {
  switch (pw.mu_family) {
    case MU_LOG: return d[0] * (1.0f - 2.0f * c);
    case MU_LEGENDRE:
    case MU_LEGENDRE_LOGPRIOR: return legendre_eval(d, pw.mu_ncoef, 2.0f * c - 1.0f);
    default: return 0.0f;
  }
}
__device__ __forceinline__ float mob_dtheta(float c, float Dval, const PointwiseParams& pw, const float* __restrict__ d) {
  switch (pw.mob_family) {
    case MOB_CONST: return d[0];
    case MOB_LEGENDRE_EXP: return Dval * legendre_eval(d, pw.mob_ncoef, 2.0f * c - 1.0f);
    default: return 0.0f;
  }
}

// pass 1: mu, D (direction 0 writes them) and mu~, D~ of every direction
static __global__ void __launch_bounds__(256) ch_tan_mu_kernel(const __grid_constant__ ChTanParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny;
  if (i >= npts) return;
  const int b = blockIdx.y, dir = blockIdx.z;
  const size_t ob = (size_t)b * npts, od = ((size_t)dir * p.batch + b) * npts;
  const int r = i / p.ny, c = i - r * p.ny;
  const int rp = (r + 1 == p.nx) ? 0 : r + 1, rm = (r == 0) ? p.nx - 1 : r - 1;
  const int cp = (c + 1 == p.ny) ? 0 : c + 1, cm = (c == 0) ? p.ny - 1 : c - 1;
  const float* u = p.u + ob;
  const float* v = p.v + od;
  const float u0 = u[i], v0 = v[i];
  const float Dv = mob<MOB_RUNTIME>(u0, p.pw);
  if (dir == 0) {
    const float lap = ((u[rp * p.ny + c] - 2.0f * u0) + u[rm * p.ny + c]) * p.inv_hx2 +
                      ((u[r * p.ny + cp] - 2.0f * u0) + u[r * p.ny + cm]) * p.inv_hy2;
    p.mu[ob + i] = mu_h<MU_RUNTIME>(u0, p.pw, 0.0f) - p.kappa * lap;
    p.dd[ob + i] = Dv;
  }
  const float lapv = ((v[rp * p.ny + c] - 2.0f * v0) + v[rm * p.ny + c]) * p.inv_hx2 +
                     ((v[r * p.ny + cp] - 2.0f * v0) + v[r * p.ny + cm]) * p.inv_hy2;
  p.mut[od + i] = mu_h_prime(u0, p.pw) * v0 + mu_h_dtheta(u0, p.pw, p.dmu + dir * PDEOPT_ADJ_NCOEF) - p.kappa * lapv;
  p.ddt[od + i] = mob_prime(u0, Dv, p.pw) * v0 + mob_dtheta(u0, Dv, p.pw, p.dmob + dir * PDEOPT_ADJ_NCOEF);
}

// pass 2: f~
static __global__ void __launch_bounds__(256) ch_tan_rhs_kernel(const __grid_constant__ ChTanParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny;
  if (i >= npts) return;
  const int b = blockIdx.y, dir = blockIdx.z;
  const size_t ob = (size_t)b * npts, od = ((size_t)dir * p.batch + b) * npts;
  const float *mu = p.mu + ob, *D = p.dd + ob, *mt = p.mut + od, *Dt = p.ddt + od;
  if (p.eq == 1) {
    p.ft[od + i] = -(Dt[i] * mu[i] + D[i] * mt[i]);
    return;
  }
  const int r = i / p.ny, c = i - r * p.ny;
  const int rp = (r + 1 == p.nx) ? 0 : r + 1, rm = (r == 0) ? p.nx - 1 : r - 1;
  const int cp = (c + 1 == p.ny) ? 0 : c + 1, cm = (c == 0) ? p.ny - 1 : c - 1;
  const float m0 = mu[i], D0 = D[i], t0 = mt[i], E0 = Dt[i];
  float acc = 0.f;
  auto dirn = [&](int ip, int im, float inv_h) {
    // face fluxes F = 1/2 (D~ + D~') (mu' - mu)/h + 1/2 (D + D') (mu~' - mu~)/h on the + and - faces
    const float Fp = 0.5f * ((E0 + Dt[ip]) * (mu[ip] - m0) + (D0 + D[ip]) * (mt[ip] - t0)) * inv_h;
    const float Fm = 0.5f * ((Dt[im] + E0) * (m0 - mu[im]) + (D[im] + D0) * (t0 - mt[im])) * inv_h;
    acc += (Fp - Fm) * inv_h;
  };
  dirn(rp * p.ny + c, rm * p.ny + c, p.inv_hx);
  dirn(r * p.ny + cp, r * p.ny + cm, p.inv_hy);
  p.ft[od + i] = acc;
}

}  // namespace pdeopt
