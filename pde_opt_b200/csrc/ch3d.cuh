// 3-D Cahn-Hilliard on a periodic grid: finite-difference RHS kernels and the line-FFT functors of
// the semi-implicit step (sm_100a).  HBM-bound streaming kernels: the 512^3 field (512 MB) does not
// fit on chip, so the design goal is the fewest passes over HBM, all of them coalesced.
//
// Replaces CahnHilliard3DPeriodic.rhs_fd (pde_opt/numerics/equations/cahn_hilliard.py:177-200) with
// the 3-D stencils of pde_opt/numerics/utils/derivatives.py:15-21, :29-36, :44-51, :59-66, and
// SemiImplicitFourierSpectral.step (solvers.py:56-70) on 3-D fields.
//
// Slab decomposition (BASELINE config 5): a rank holds nx planes [nx][ny][nz] of the global field;
// the two planes on either side come in through halo arrays (null = periodic wrap on this rank).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "linefft.cuh"
#include "pointwise.cuh"

namespace pdeopt {

struct Ch3dParams {
  int nx, ny, nz, batch;
  const float* u;        // [batch][nx][ny][nz]
  const float* halo_lo;  // [2][ny][nz]: planes x = -2, -1 (or null)
  const float* halo_hi;  // [2][ny][nz]: planes x = nx, nx+1 (or null)
  float* mu;             // [nx+2][ny][nz]: planes x = -1 .. nx
  float* f;              // [nx][ny][nz]
  float inv_hx, inv_hy, inv_hz, inv_hx2, inv_hy2, inv_hz2, kappa;
  PointwiseParams pw;
};

// plane x of domain b; halos (slab mode) are only meaningful for a single domain
__device__ __forceinline__ const float* ch3d_plane(const Ch3dParams& p, int b, int x) {
  const size_t pl = (size_t)p.ny * p.nz;
  const float* u = p.u + (size_t)b * p.nx * pl;
  if (x < 0) return p.halo_lo ? p.halo_lo + (size_t)(x + 2) * pl : u + (size_t)(x + p.nx) * pl;
  if (x >= p.nx) return p.halo_hi ? p.halo_hi + (size_t)(x - p.nx) * pl : u + (size_t)(x - p.nx) * pl;
  return u + (size_t)x * pl;
}

// mu = mu_h(u) - kappa * lap(u) on planes x = -1 .. nx   (cahn_hilliard.py:179; derivatives.py:15-21)
__global__ void __launch_bounds__(256) ch3d_mu_kernel(const __grid_constant__ Ch3dParams p) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z / (p.nx + 2);
  const int xm = blockIdx.z % (p.nx + 2);  // 0 .. nx+1  <->  x = xm - 1
  if (z >= p.nz) return;
  const int x = xm - 1;
  const float* c0 = ch3d_plane(p, b, x);
  const float* cm = ch3d_plane(p, b, x - 1);
  const float* cp = ch3d_plane(p, b, x + 1);
  const int yp = (y + 1 == p.ny) ? 0 : y + 1, ym = (y == 0) ? p.ny - 1 : y - 1;
  const int zp = (z + 1 == p.nz) ? 0 : z + 1, zm = (z == 0) ? p.nz - 1 : z - 1;
  const size_t o = (size_t)y * p.nz + z;
  const float u = c0[o];
  const float lap = ((cp[o] - 2.0f * u) + cm[o]) * p.inv_hx2 +
                    ((c0[(size_t)yp * p.nz + z] - 2.0f * u) + c0[(size_t)ym * p.nz + z]) * p.inv_hy2 +
                    ((c0[(size_t)y * p.nz + zp] - 2.0f * u) + c0[(size_t)y * p.nz + zm]) * p.inv_hz2;
  p.mu[((size_t)b * (p.nx + 2) + xm) * p.ny * p.nz + o] = mu_h<MU_RUNTIME>(u, p.pw, 0.0f) - p.kappa * lap;
}

// f = div( avg_face(D(u)) * grad_face(mu) )   (cahn_hilliard.py:181-200)
__global__ void __launch_bounds__(256) ch3d_div_kernel(const __grid_constant__ Ch3dParams p) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z / p.nx, x = blockIdx.z % p.nx;
  if (z >= p.nz) return;
  const size_t pl = (size_t)p.ny * p.nz;
  const float* u0 = ch3d_plane(p, b, x);
  const float* um = ch3d_plane(p, b, x - 1);
  const float* up = ch3d_plane(p, b, x + 1);
  const float* m0 = p.mu + ((size_t)b * (p.nx + 2) + x + 1) * pl;
  const float* mm = m0 - pl;
  const float* mp = m0 + pl;
  const int yp = (y + 1 == p.ny) ? 0 : y + 1, ym = (y == 0) ? p.ny - 1 : y - 1;
  const int zp = (z + 1 == p.nz) ? 0 : z + 1, zm = (z == 0) ? p.nz - 1 : z - 1;
  const size_t o = (size_t)y * p.nz + z;
  const size_t oyp = (size_t)yp * p.nz + z, oym = (size_t)ym * p.nz + z;
  const size_t ozp = (size_t)y * p.nz + zp, ozm = (size_t)y * p.nz + zm;
  const float D0 = mob<MOB_RUNTIME>(u0[o], p.pw), mu0 = m0[o];
  auto flux = [&](float Dn, float mun, float inv_h) { return (0.5f * (D0 + Dn)) * ((mun - mu0) * inv_h); };
  auto fluxb = [&](float Dn, float mun, float inv_h) { return (0.5f * (Dn + D0)) * ((mu0 - mun) * inv_h); };
  const float fx = (flux(mob<MOB_RUNTIME>(up[o], p.pw), mp[o], p.inv_hx) - fluxb(mob<MOB_RUNTIME>(um[o], p.pw), mm[o], p.inv_hx)) * p.inv_hx;
  const float fy = (flux(mob<MOB_RUNTIME>(u0[oyp], p.pw), m0[oyp], p.inv_hy) - fluxb(mob<MOB_RUNTIME>(u0[oym], p.pw), m0[oym], p.inv_hy)) * p.inv_hy;
  const float fz = (flux(mob<MOB_RUNTIME>(u0[ozp], p.pw), m0[ozp], p.inv_hz) - fluxb(mob<MOB_RUNTIME>(u0[ozm], p.pw), m0[ozm], p.inv_hz)) * p.inv_hz;
  p.f[((size_t)b * p.nx + x) * pl + o] = (fx + fy) + fz;
}

// ---- line-FFT functors of the 3-D semi-implicit step -------------------------------------------
struct LfLoadReal {  // real array -> complex with zero imaginary part
  const float* p;
  LineGeom g;
  __device__ __forceinline__ LineGeom gin() const { return g; }
  __device__ __forceinline__ float2 load(long long off, long long, int) const { return make_float2(p[off], 0.f); }
};
// multiplier scale / (1 + dt * symbol) with the (A-folded, position-ordered) symbol laid out like the data
struct LfMidImex {
  const float* sym;
  LineGeom g;
  float dt, scale;
  __device__ __forceinline__ LineGeom gaux() const { return g; }
  __device__ __forceinline__ float2 apply(float2 v, long long off_aux, long long, int) const {
    const float m = __fdividef(scale, fmaf(dt, sym[off_aux], 1.0f));
    return make_float2(v.x * m, v.y * m);
  }
};
// y1 = y0 + dt * Re(g)   (solvers.py:63)
struct LfStoreUpdate {
  const float* y0;
  float* y1;
  LineGeom g;
  float dt;
  __device__ __forceinline__ LineGeom gout() const { return g; }
  __device__ __forceinline__ void store(long long off, long long, int, float2 v) const { y1[off] = fmaf(dt, v.x, y0[off]); }
  __device__ __forceinline__ void flush(long long) {}
};

}  // namespace pdeopt
