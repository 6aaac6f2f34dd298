// Finite-difference phase-field right-hand sides with an EXTERNALLY evaluated homogeneous chemical potential
// (and optionally mobility): the unfused-but-batched path for mu / D closures that are not pointwise families —
// the reference's PeriodicCNN and Mixer2d networks (pde_opt/numerics/functions/cnn.py:46-102, mixer_mlp.py:40-86,
// docs/notebooks/optimization_neural_network.ipynb) or any other callable.  The closure runs in the caller's
// framework on the whole batch; the stencils of cahn_hilliard.py:89-109 / allen_cahn.py:81-84
// (utils/derivatives.py:8-66) run here, and the result feeds pdeopt_sifs_filter_batched (solvers.py:62-63).
//   pass 1: mu = mu_h - kappa lap(u),  D = given field or the plan's enumerated family
//   pass 2: CH f = div( D_face grad_face(mu) );  AC: f = -D mu (done in pass 1)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pointwise.cuh"

namespace pdeopt {

struct GivenMuParams {
  int nx, ny, batch, eq;  // eq: 0 = Cahn-Hilliard, 1 = Allen-Cahn
  int keep_mu;            // adjoint use: always write mu and D (Allen-Cahn normally finishes in pass 1)
  const float* u;    // [B][nx][ny]
  const float* muh;  // [B][nx][ny] mu_h(u) evaluated by the caller
  const float* mob;  // [B][nx][ny] D(u) / R(u) evaluated by the caller, or null: the plan's family
  float* mu;         // [B][nx][ny] scratch
  float* dd;         // [B][nx][ny] scratch
  float* f;          // [B][nx][ny] out
  float inv_hx, inv_hy, inv_hx2, inv_hy2, kappa;
  PointwiseParams pw;
};

static __global__ void __launch_bounds__(256) given_mu_pass1_kernel(const __grid_constant__ GivenMuParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny;
  if (i >= npts) return;
  const size_t o = (size_t)blockIdx.y * npts;
  const int r = i / p.ny, c = i - r * p.ny;
  const int rp = (r + 1 == p.nx) ? 0 : r + 1, rm = (r == 0) ? p.nx - 1 : r - 1;
  const int cp = (c + 1 == p.ny) ? 0 : c + 1, cm = (c == 0) ? p.ny - 1 : c - 1;
  const float* u = p.u + o;
  const float u0 = u[i];
  const float lap = ((u[rp * p.ny + c] - 2.0f * u0) + u[rm * p.ny + c]) * p.inv_hx2 +
                    ((u[r * p.ny + cp] - 2.0f * u0) + u[r * p.ny + cm]) * p.inv_hy2;
  const float mu = p.muh[o + i] - p.kappa * lap;
  const float D = p.mob ? p.mob[o + i] : mob<MOB_RUNTIME>(u0, p.pw);
  if (p.eq == 1 && !p.keep_mu) {
    p.f[o + i] = -D * mu;  // allen_cahn.py:84
  } else {
    p.mu[o + i] = mu;
    p.dd[o + i] = D;
  }
}

// adjoint helper: out = a - kappa lap(b)
static __global__ void __launch_bounds__(256) sub_kappa_lap_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                   float* __restrict__ out, int nx, int ny, float inv_hx2, float inv_hy2,
                                                                   float kappa) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = nx * ny;
  if (i >= npts) return;
  const size_t o = (size_t)blockIdx.y * npts;
  const int r = i / ny, c = i - r * ny;
  const int rp = (r + 1 == nx) ? 0 : r + 1, rm = (r == 0) ? nx - 1 : r - 1;
  const int cp = (c + 1 == ny) ? 0 : c + 1, cm = (c == 0) ? ny - 1 : c - 1;
  const float* bb = b + o;
  const float b0 = bb[i];
  const float lap = ((bb[rp * ny + c] - 2.0f * b0) + bb[rm * ny + c]) * inv_hx2 + ((bb[r * ny + cp] - 2.0f * b0) + bb[r * ny + cm]) * inv_hy2;
  out[o + i] = a[o + i] - kappa * lap;
}

static __global__ void __launch_bounds__(256) given_mu_pass2_kernel(const __grid_constant__ GivenMuParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny;
  if (i >= npts) return;
  const size_t o = (size_t)blockIdx.y * npts;
  const int r = i / p.ny, c = i - r * p.ny;
  const int rp = (r + 1 == p.nx) ? 0 : r + 1, rm = (r == 0) ? p.nx - 1 : r - 1;
  const int cp = (c + 1 == p.ny) ? 0 : c + 1, cm = (c == 0) ? p.ny - 1 : c - 1;
  const float *mu = p.mu + o, *D = p.dd + o;
  const float m0 = mu[i], D0 = D[i];
  // cahn_hilliard.py:100-109: face mobility = average, face gradient = forward difference, divergence = backward difference
  const float Fxp = 0.5f * (D0 + D[rp * p.ny + c]) * ((mu[rp * p.ny + c] - m0) * p.inv_hx);
  const float Fxm = 0.5f * (D[rm * p.ny + c] + D0) * ((m0 - mu[rm * p.ny + c]) * p.inv_hx);
  const float Fyp = 0.5f * (D0 + D[r * p.ny + cp]) * ((mu[r * p.ny + cp] - m0) * p.inv_hy);
  const float Fym = 0.5f * (D[r * p.ny + cm] + D0) * ((m0 - mu[r * p.ny + cm]) * p.inv_hy);
  p.f[o + i] = (Fxp - Fxm) * p.inv_hx + (Fyp - Fym) * p.inv_hy;
}

}  // namespace pdeopt
