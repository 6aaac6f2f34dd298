// Discrete adjoint of one semi-implicit step of the finite-difference phase-field equations
// (Cahn-Hilliard / Allen-Cahn 2-D), with the cotangents of the closure coefficients reduced in-kernel.
//
// Forward step (solvers.py:56-70 with cahn_hilliard.py:89-109 / allen_cahn.py:81-84):
//     y1 = u + dt G f(u),   G = F^-1 [1/(1 + A dt sigma)] F  (self-adjoint),
//     CH: f = div( D_face grad_face(mu) ),  AC: f = -R(u) mu,   mu = mu_h(u; theta) - kappa lap(u).
// Backward, with w = dt G lam1 (computed by the fused filter kernel, pdeopt_sifs_filter_batched):
//     CH:  mu_bar = div( D_face grad_face(w) )                       (the flux operator is symmetric in mu <-> w)
//          D_bar  = -1/2 sum_faces grad_face(w) grad_face(mu)        (faces adjacent to the cell)
//     AC:  mu_bar = -R(u) w,   D_bar (= R_bar) = -w mu
//     lam0 = lam1 + mu_h'(u) mu_bar - kappa lap(mu_bar) + D'(u) D_bar
//     dL/dtheta_mu += sum  d mu_h / d theta (u) mu_bar,   dL/dtheta_D += sum  d D / d theta (u) D_bar.
// This is what reverse-mode differentiation through diffeqsolve produces for PDEModel.mse
// (pde_model.py:274-323) when the optimised leaves are the Legendre coefficients of mu and D
// (docs/notebooks/optimization_3D.ipynb).  Streaming global-memory kernels: the adjoint is used in
// training loops with few trajectories, where launch count, not bandwidth, is what matters.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pointwise.cuh"

#ifndef PDEOPT_ADJ_NCOEF
#define PDEOPT_ADJ_NCOEF 16  // == PDEOPT_MAX_COEF (include/pdeopt_b200.h)
#endif

namespace pdeopt {

struct ChAdjParams {
  int nx, ny, batch, eq;  // eq: 0 = Cahn-Hilliard, 1 = Allen-Cahn
  const float* u;     // [B][nx][ny] state at the start of the step
  const float* w;     // [B][nx][ny] dt * G lam1
  const float* lam1;  // [B][nx][ny]
  float* lam0;        // [B][nx][ny]
  float* mu;          // [B][nx][ny] scratch
  float* dd;          // [B][nx][ny] scratch: D(u)
  float* mub;         // [B][nx][ny] scratch: mu_bar
  float* db;          // [B][nx][ny] scratch: D_bar
  double* gmu;        // [B][16] accumulated cotangents of mu_coef (float64: thousands of steps of signed terms)
  double* gmob;       // [B][16] accumulated cotangents of mob_coef
  float inv_hx, inv_hy, inv_hx2, inv_hy2, kappa;
  PointwiseParams pw;
};

// Legendre values P_n(x) and the expansion's derivative d/dx sum a_n P_n(x)
__device__ __forceinline__ float legendre_deriv(const float* __restrict__ coef, int ncoef, float x) {
  // P_n'(x) by the recurrence P_n' = P_{n-2}' + (2n - 1) P_{n-1}
  float d = 0.f;
  float p_prev = 1.0f, p_curr = x;      // P_0, P_1
  float dp_prev = 0.0f, dp_curr = 1.0f; // P_0', P_1'
  if (ncoef > 1) d = coef[1];
  static_for<2, 16>([&](auto nc) {
    constexpr int n = decltype(nc)::value;
    if (n < ncoef) {
      const float p_next = LegC<n>::a * x * p_curr - LegC<n>::b * p_prev;
      const float dp_next = dp_prev + float(2 * n - 1) * p_curr;
      d = fmaf(coef[n], dp_next, d);
      p_prev = p_curr;
      p_curr = p_next;
      dp_prev = dp_curr;
      dp_curr = dp_next;
    }
  });
  return d;
}

__device__ __forceinline__ float mu_h_prime(float c, const PointwiseParams& pw) {
  switch (pw.mu_family) {
    case MU_DOUBLE_WELL: return 3.0f * c * c - 1.0f;
    case MU_LOG: return 1.0f / (c * (1.0f - c)) - 2.0f * pw.mu_coef[0];
    case MU_LEGENDRE: return 2.0f * legendre_deriv(pw.mu_coef, pw.mu_ncoef, 2.0f * c - 1.0f);
    default: return 2.0f * legendre_deriv(pw.mu_coef, pw.mu_ncoef, 2.0f * c - 1.0f) + 1.0f / (c * (1.0f - c));
  }
}
__device__ __forceinline__ float mob_prime(float c, float Dval, const PointwiseParams& pw) {
  switch (pw.mob_family) {
    case MOB_CONST: return 0.0f;
    case MOB_DEGENERATE: return 1.0f - 2.0f * c;
    case MOB_ONE_PLUS_SQ: return 2.0f * c;
    default: return Dval * 2.0f * legendre_deriv(pw.mob_coef, pw.mob_ncoef, 2.0f * c - 1.0f);
  }
}

// pass 1: mu and D of the saved state
static __global__ void __launch_bounds__(256) ch_adj_mu_kernel(const __grid_constant__ ChAdjParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny;
  if (i >= npts) return;
  const int b = blockIdx.y;
  const int r = i / p.ny, c = i - r * p.ny;
  const float* u = p.u + (size_t)b * npts;
  const int rp = (r + 1 == p.nx) ? 0 : r + 1, rm = (r == 0) ? p.nx - 1 : r - 1;
  const int cp = (c + 1 == p.ny) ? 0 : c + 1, cm = (c == 0) ? p.ny - 1 : c - 1;
  const float u0 = u[i];
  const float lap = ((u[rp * p.ny + c] - 2.0f * u0) + u[rm * p.ny + c]) * p.inv_hx2 +
                    ((u[r * p.ny + cp] - 2.0f * u0) + u[r * p.ny + cm]) * p.inv_hy2;
  p.mu[(size_t)b * npts + i] = mu_h<MU_RUNTIME>(u0, p.pw, 0.0f) - p.kappa * lap;
  p.dd[(size_t)b * npts + i] = mob<MOB_RUNTIME>(u0, p.pw);
}

// pass 2: mu_bar and D_bar
static __global__ void __launch_bounds__(256) ch_adj_bar_kernel(const __grid_constant__ ChAdjParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny;
  if (i >= npts) return;
  const int b = blockIdx.y;
  const size_t o = (size_t)b * npts;
  const int r = i / p.ny, c = i - r * p.ny;
  const float* mu = p.mu + o;
  const float* D = p.dd + o;
  const float* w = p.w + o;
  if (p.eq == 1) {
    p.mub[o + i] = -D[i] * w[i];
    p.db[o + i] = -w[i] * mu[i];
    return;
  }
  const int rp = (r + 1 == p.nx) ? 0 : r + 1, rm = (r == 0) ? p.nx - 1 : r - 1;
  const int cp = (c + 1 == p.ny) ? 0 : c + 1, cm = (c == 0) ? p.ny - 1 : c - 1;
  const int ixp = rp * p.ny + c, ixm = rm * p.ny + c, iyp = r * p.ny + cp, iym = r * p.ny + cm;
  const float w0 = w[i], m0 = mu[i], D0 = D[i];
  const float gwxp = (w[ixp] - w0) * p.inv_hx, gwxm = (w0 - w[ixm]) * p.inv_hx;
  const float gwyp = (w[iyp] - w0) * p.inv_hy, gwym = (w0 - w[iym]) * p.inv_hy;
  const float gmxp = (mu[ixp] - m0) * p.inv_hx, gmxm = (m0 - mu[ixm]) * p.inv_hx;
  const float gmyp = (mu[iyp] - m0) * p.inv_hy, gmym = (m0 - mu[iym]) * p.inv_hy;
  const float Dxp = 0.5f * (D0 + D[ixp]), Dxm = 0.5f * (D[ixm] + D0);
  const float Dyp = 0.5f * (D0 + D[iyp]), Dym = 0.5f * (D[iym] + D0);
  p.mub[o + i] = (Dxp * gwxp - Dxm * gwxm) * p.inv_hx + (Dyp * gwyp - Dym * gwym) * p.inv_hy;
  p.db[o + i] = -0.5f * ((gwxp * gmxp + gwxm * gmxm) + (gwyp * gmyp + gwym * gmym));
}

// pass 3: lam0 and the coefficient cotangents (block reduction + one atomic per block and coefficient)
static __global__ void __launch_bounds__(256) ch_adj_out_kernel(const __grid_constant__ ChAdjParams p) {
  __shared__ float red[8][33];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny;
  const int b = blockIdx.y;
  const size_t o = (size_t)b * npts;
  float gm[PDEOPT_ADJ_NCOEF], gd[PDEOPT_ADJ_NCOEF];
#pragma unroll
  for (int n = 0; n < PDEOPT_ADJ_NCOEF; ++n) gm[n] = gd[n] = 0.f;
  if (i < npts) {
    const int r = i / p.ny, c = i - r * p.ny;
    const int rp = (r + 1 == p.nx) ? 0 : r + 1, rm = (r == 0) ? p.nx - 1 : r - 1;
    const int cp = (c + 1 == p.ny) ? 0 : c + 1, cm = (c == 0) ? p.ny - 1 : c - 1;
    const float* mb = p.mub + o;
    const float u0 = p.u[o + i], mb0 = mb[i], db0 = p.db[o + i], D0 = p.dd[o + i];
    const float lap = ((mb[rp * p.ny + c] - 2.0f * mb0) + mb[rm * p.ny + c]) * p.inv_hx2 +
                      ((mb[r * p.ny + cp] - 2.0f * mb0) + mb[r * p.ny + cm]) * p.inv_hy2;
    p.lam0[o + i] = p.lam1[o + i] + (mu_h_prime(u0, p.pw) * mb0 - p.kappa * lap) + mob_prime(u0, D0, p.pw) * db0;
    // d mu_h / d theta and d D / d theta
    const float x = 2.0f * u0 - 1.0f;
    if (p.pw.mu_family == MU_LOG) {
      gm[0] = (1.0f - 2.0f * u0) * mb0;  // mu_h = logit(c) + w (1 - 2c)
    } else if (p.pw.mu_family == MU_LEGENDRE || p.pw.mu_family == MU_LEGENDRE_LOGPRIOR) {
      float pp = 1.0f, pc = x;
      gm[0] = mb0;
      if (p.pw.mu_ncoef > 1) gm[1] = x * mb0;
#pragma unroll
      for (int n = 2; n < PDEOPT_ADJ_NCOEF; ++n) {
        if (n < p.pw.mu_ncoef) {
          const float pn = (float(2 * n - 1) * x * pc - float(n - 1) * pp) / float(n);
          gm[n] = pn * mb0;
          pp = pc;
          pc = pn;
        }
      }
    }
    if (p.pw.mob_family == MOB_CONST) {
      gd[0] = db0;
    } else if (p.pw.mob_family == MOB_LEGENDRE_EXP) {
      float pp = 1.0f, pc = x;
      gd[0] = D0 * db0;
      if (p.pw.mob_ncoef > 1) gd[1] = D0 * x * db0;
#pragma unroll
      for (int n = 2; n < PDEOPT_ADJ_NCOEF; ++n) {
        if (n < p.pw.mob_ncoef) {
          const float pn = (float(2 * n - 1) * x * pc - float(n - 1) * pp) / float(n);
          gd[n] = D0 * pn * db0;
          pp = pc;
          pc = pn;
        }
      }
    }
  }
  // reduce 2 * NCOEF values over the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int n = 0; n < 2 * PDEOPT_ADJ_NCOEF; ++n) {
    float v = n < PDEOPT_ADJ_NCOEF ? gm[n] : gd[n - PDEOPT_ADJ_NCOEF];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) red[warp][n] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2 * PDEOPT_ADJ_NCOEF) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
    if (t != 0.f) {
      if (threadIdx.x < PDEOPT_ADJ_NCOEF) atomicAdd(p.gmu + (size_t)b * PDEOPT_ADJ_NCOEF + threadIdx.x, (double)t);
      else atomicAdd(p.gmob + (size_t)b * PDEOPT_ADJ_NCOEF + threadIdx.x - PDEOPT_ADJ_NCOEF, (double)t);
    }
  }
}

}  // namespace pdeopt

// ---- 3-D forms (CahnHilliard3DPeriodic, cahn_hilliard.py:177-200): same three passes --------------
namespace pdeopt {

struct Ch3AdjParams {
  int nx, ny, nz, batch;
  const float *u, *w, *lam1;
  float *lam0, *mu, *dd, *mub, *db;
  double *gmu, *gmob;
  float inv_hx, inv_hy, inv_hz, inv_hx2, inv_hy2, inv_hz2, kappa;
  PointwiseParams pw;
};

struct Ch3Idx {
  int i, ixp, ixm, iyp, iym, izp, izm;
  __device__ __forceinline__ Ch3Idx(const Ch3AdjParams& p, int idx) {
    const int pl = p.ny * p.nz;
    const int x = idx / pl, rem = idx - x * pl, y = rem / p.nz, z = rem - y * p.nz;
    i = idx;
    ixp = ((x + 1 == p.nx) ? 0 : x + 1) * pl + rem;
    ixm = ((x == 0) ? p.nx - 1 : x - 1) * pl + rem;
    iyp = x * pl + ((y + 1 == p.ny) ? 0 : y + 1) * p.nz + z;
    iym = x * pl + ((y == 0) ? p.ny - 1 : y - 1) * p.nz + z;
    izp = x * pl + y * p.nz + ((z + 1 == p.nz) ? 0 : z + 1);
    izm = x * pl + y * p.nz + ((z == 0) ? p.nz - 1 : z - 1);
  }
};

static __global__ void __launch_bounds__(256) ch3_adj_mu_kernel(const __grid_constant__ Ch3AdjParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny * p.nz;
  if (idx >= npts) return;
  const size_t o = (size_t)blockIdx.y * npts;
  const float* u = p.u + o;
  const Ch3Idx n(p, idx);
  const float u0 = u[idx];
  const float lap = ((u[n.ixp] - 2.0f * u0) + u[n.ixm]) * p.inv_hx2 + ((u[n.iyp] - 2.0f * u0) + u[n.iym]) * p.inv_hy2 +
                    ((u[n.izp] - 2.0f * u0) + u[n.izm]) * p.inv_hz2;
  p.mu[o + idx] = mu_h<MU_RUNTIME>(u0, p.pw, 0.0f) - p.kappa * lap;
  p.dd[o + idx] = mob<MOB_RUNTIME>(u0, p.pw);
}

static __global__ void __launch_bounds__(256) ch3_adj_bar_kernel(const __grid_constant__ Ch3AdjParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny * p.nz;
  if (idx >= npts) return;
  const size_t o = (size_t)blockIdx.y * npts;
  const float *mu = p.mu + o, *D = p.dd + o, *w = p.w + o;
  const Ch3Idx n(p, idx);
  const float w0 = w[idx], m0 = mu[idx], D0 = D[idx];
  float mub = 0.f, db = 0.f;
  auto dir = [&](int ip, int im, float inv_h) {
    const float gwp = (w[ip] - w0) * inv_h, gwm = (w0 - w[im]) * inv_h;
    const float gmp = (mu[ip] - m0) * inv_h, gmm = (m0 - mu[im]) * inv_h;
    mub += (0.5f * (D0 + D[ip]) * gwp - 0.5f * (D[im] + D0) * gwm) * inv_h;
    db += gwp * gmp + gwm * gmm;
  };
  dir(n.ixp, n.ixm, p.inv_hx);
  dir(n.iyp, n.iym, p.inv_hy);
  dir(n.izp, n.izm, p.inv_hz);
  p.mub[o + idx] = mub;
  p.db[o + idx] = -0.5f * db;
}

static __global__ void __launch_bounds__(256) ch3_adj_out_kernel(const __grid_constant__ Ch3AdjParams p) {
  __shared__ float red[8][33];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int npts = p.nx * p.ny * p.nz;
  const int b = blockIdx.y;
  const size_t o = (size_t)b * npts;
  float gm[PDEOPT_ADJ_NCOEF], gd[PDEOPT_ADJ_NCOEF];
#pragma unroll
  for (int n = 0; n < PDEOPT_ADJ_NCOEF; ++n) gm[n] = gd[n] = 0.f;
  if (idx < npts) {
    const Ch3Idx n(p, idx);
    const float* mb = p.mub + o;
    const float u0 = p.u[o + idx], mb0 = mb[idx], db0 = p.db[o + idx], D0 = p.dd[o + idx];
    const float lap = ((mb[n.ixp] - 2.0f * mb0) + mb[n.ixm]) * p.inv_hx2 + ((mb[n.iyp] - 2.0f * mb0) + mb[n.iym]) * p.inv_hy2 +
                      ((mb[n.izp] - 2.0f * mb0) + mb[n.izm]) * p.inv_hz2;
    p.lam0[o + idx] = p.lam1[o + idx] + (mu_h_prime(u0, p.pw) * mb0 - p.kappa * lap) + mob_prime(u0, D0, p.pw) * db0;
    const float x = 2.0f * u0 - 1.0f;
    if (p.pw.mu_family == MU_LOG) {
      gm[0] = (1.0f - 2.0f * u0) * mb0;
    } else if (p.pw.mu_family == MU_LEGENDRE || p.pw.mu_family == MU_LEGENDRE_LOGPRIOR) {
      float pp = 1.0f, pc = x;
      gm[0] = mb0;
      if (p.pw.mu_ncoef > 1) gm[1] = x * mb0;
#pragma unroll
      for (int k = 2; k < PDEOPT_ADJ_NCOEF; ++k) {
        if (k < p.pw.mu_ncoef) {
          const float pn = (float(2 * k - 1) * x * pc - float(k - 1) * pp) / float(k);
          gm[k] = pn * mb0;
          pp = pc;
          pc = pn;
        }
      }
    }
    if (p.pw.mob_family == MOB_CONST) {
      gd[0] = db0;
    } else if (p.pw.mob_family == MOB_LEGENDRE_EXP) {
      float pp = 1.0f, pc = x;
      gd[0] = D0 * db0;
      if (p.pw.mob_ncoef > 1) gd[1] = D0 * x * db0;
#pragma unroll
      for (int k = 2; k < PDEOPT_ADJ_NCOEF; ++k) {
        if (k < p.pw.mob_ncoef) {
          const float pn = (float(2 * k - 1) * x * pc - float(k - 1) * pp) / float(k);
          gd[k] = D0 * pn * db0;
          pp = pc;
          pc = pn;
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 2 * PDEOPT_ADJ_NCOEF; ++k) {
    float v = k < PDEOPT_ADJ_NCOEF ? gm[k] : gd[k - PDEOPT_ADJ_NCOEF];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2 * PDEOPT_ADJ_NCOEF) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
    if (t != 0.f) {
      if (threadIdx.x < PDEOPT_ADJ_NCOEF) atomicAdd(p.gmu + (size_t)b * PDEOPT_ADJ_NCOEF + threadIdx.x, (double)t);
      else atomicAdd(p.gmob + (size_t)b * PDEOPT_ADJ_NCOEF + threadIdx.x - PDEOPT_ADJ_NCOEF, (double)t);
    }
  }
}

}  // namespace pdeopt
