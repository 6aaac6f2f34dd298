// Generic-size fused advection-diffusion stepper (forward only): any power-of-two nx, ny with
// nx*ny <= 8192.  Same arithmetic as ad128.cuh — two environments per complex field, odd multipliers
// zeroed on the Nyquist lines, 3 forward + 1 inverse transform per step — with the simple radix-2
// shared-memory stages of sifs_generic.cuh.  It serves the sizes the tuned 128x128 kernel does not,
// in particular the reference's own 64x64 advection-diffusion run whose final state survives as
// notebooks/reference.npy (the fixture tests/golden/ref_advection_diffusion_64.npy).
//
// Replaces SemiImplicitFourierSpectral.step (pde_opt/numerics/solvers.py:56-70) with the recovered
// AdvectionDiffusion2D.rhs (notebooks/run_advection_diffusion.ipynb cells 0-2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ad128.cuh"
#include "sifs_generic.cuh"

namespace pdeopt {

struct AdGenParams {
  AdParams a;
  int nx, ny, lognx, logny;
};

inline size_t ad_gen_smem_bytes(int nx, int ny) {
  return sizeof(float2) * (size_t)(3 * nx * ny + 3 * nx + 3 * ny + 8);
}

__global__ void __launch_bounds__(kGenThreads, 1) ad_generic_fwd_kernel(const __grid_constant__ AdGenParams gp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const AdParams& p = gp.a;
  const int nx = gp.nx, ny = gp.ny, npts = nx * ny, tlx = nx / 2 + 1, tly = ny / 2 + 1;
  float2* U = reinterpret_cast<float2*>(smem_raw);
  float2* Z = U + npts;
  float2* ACC = Z + npts;
  float2* twx = ACC + npts;
  float2* twy = twx + nx;
  float2* ax = twy + ny;   // -(p0/p1) dx ex   (row tables)
  float2* ex = ax + nx;
  float2* ay = ex + nx;    // column tables
  float2* ey = ay + ny;
  const int tid = threadIdx.x;
  const int env_a = 2 * blockIdx.x;
  const int env_b = (env_a + 1 < p.batch) ? env_a + 1 : env_a;
  const bool b_valid = env_a + 1 < p.batch;
  const float* tabA = p.tabA;
  const float* tabL = p.tabA + tlx * tly;
  const float* kxs = tabL + tlx * tly;
  const float* kys = kxs + nx;

  for (int i = tid; i < nx; i += kGenThreads) {
    float s, c;
    sincospif(-2.0f * float(i) / float(nx), &s, &c);
    twx[i] = make_float2(c, s);
  }
  for (int i = tid; i < ny; i += kGenThreads) {
    float s, c;
    sincospif(-2.0f * float(i) / float(ny), &s, &c);
    twy[i] = make_float2(c, s);
  }
  {
    const float* ya = p.y0 + (size_t)env_a * npts;
    const float* yb = p.y0 + (size_t)env_b * npts;
    for (int i = tid; i < npts; i += kGenThreads) U[i] = make_float2(ya[i], yb[i]);
  }
  auto fft_fwd = [&]() {
    for (int s = 0; s < gp.logny; ++s) {
      r2_stage<false>(Z, twy, s, ny >> (s + 1), gp.logny, 1, nx, ny, npts);
      __syncthreads();
    }
    for (int s = 0; s < gp.lognx; ++s) {
      r2_stage<false>(Z, twx, s, nx >> (s + 1), gp.lognx, ny, ny, 1, npts);
      __syncthreads();
    }
  };
  int cur_seg = -1;
  for (int k = 0; k < p.ksteps; ++k) {
    const int seg = ad_seg(p, k);
    __syncthreads();
    if (seg != cur_seg) {
      const float* ca = p.ctrl + ((size_t)env_a * p.nseg + seg) * kAdCtrl;
      const float* cb = p.ctrl + ((size_t)env_b * p.nseg + seg) * kAdCtrl;
      for (int i = tid; i < nx + ny; i += kGenThreads) {
        const bool isx = i < nx;
        const int q = isx ? i : i - nx;
        const float pos = isx ? (p.lo_x + (q + 0.5f) * p.hx) : (p.lo_y + (q + 0.5f) * p.hy);
        const float2 d = make_float2(pos - (isx ? ca[0] : ca[1]), pos - (isx ? cb[0] : cb[1]));
        const float2 e = make_float2(expf(-d.x * d.x / (2.0f * ca[3])), expf(-d.y * d.y / (2.0f * cb[3])));
        const float2 a = make_float2(ca[2] * (-d.x / ca[3] * e.x), cb[2] * (-d.y / cb[3] * e.y));
        if (isx) { ax[q] = a; ex[q] = e; } else { ay[q] = a; ey[q] = e; }
      }
      cur_seg = seg;
      __syncthreads();
    }
    const float dt = p.dt[k];
    const float inv_n = 1.0f / float(npts);
    // ---- ACC = -L F[u] ----
    for (int i = tid; i < npts; i += kGenThreads) Z[i] = U[i];
    __syncthreads();
    fft_fwd();
    for (int i = tid; i < npts; i += kGenThreads) {
      const int r = i >> gp.logny, c = i & (ny - 1);
      const int kx = brev_rt(r, gp.lognx), ky = brev_rt(c, gp.logny);
      const int fx = kx <= nx / 2 ? kx : nx - kx, fy = ky <= ny / 2 ? ky : ny - ky;
      const float l = -tabL[fx * tly + fy];
      ACC[i] = make_float2(Z[i].x * l, Z[i].y * l);
    }
    __syncthreads();
    // ---- ACC += -i kx F[vx u] ----
    for (int i = tid; i < npts; i += kGenThreads) {
      const int r = i >> gp.logny, c = i & (ny - 1);
      Z[i] = f2mul(f2mul(U[i], ax[r]), ey[c]);
    }
    __syncthreads();
    fft_fwd();
    for (int i = tid; i < npts; i += kGenThreads) {
      const float kk = kxs[brev_rt(i >> gp.logny, gp.lognx)];
      ACC[i] = make_float2(fmaf(Z[i].y, kk, ACC[i].x), fmaf(-Z[i].x, kk, ACC[i].y));
    }
    __syncthreads();
    // ---- ACC += -i ky F[vy u];  filter ----
    for (int i = tid; i < npts; i += kGenThreads) {
      const int r = i >> gp.logny, c = i & (ny - 1);
      Z[i] = f2mul(f2mul(U[i], ex[r]), ay[c]);
    }
    __syncthreads();
    fft_fwd();
    for (int i = tid; i < npts; i += kGenThreads) {
      const int r = i >> gp.logny, c = i & (ny - 1);
      const int kx = brev_rt(r, gp.lognx), ky = brev_rt(c, gp.logny);
      const int fx = kx <= nx / 2 ? kx : nx - kx, fy = ky <= ny / 2 ? ky : ny - ky;
      const float kk = kys[ky];
      const float m = __fdividef(inv_n, fmaf(dt, tabA[fx * tly + fy], 1.0f));
      Z[i] = make_float2(fmaf(Z[i].y, kk, ACC[i].x) * m, fmaf(-Z[i].x, kk, ACC[i].y) * m);
    }
    __syncthreads();
    for (int s = gp.lognx - 1; s >= 0; --s) {
      r2_stage<true>(Z, twx, s, nx >> (s + 1), gp.lognx, ny, ny, 1, npts);
      __syncthreads();
    }
    for (int s = gp.logny - 1; s >= 0; --s) {
      r2_stage<true>(Z, twy, s, ny >> (s + 1), gp.logny, 1, nx, ny, npts);
      __syncthreads();
    }
    for (int i = tid; i < npts; i += kGenThreads) {
      U[i].x = fmaf(dt, Z[i].x, U[i].x);
      U[i].y = fmaf(dt, Z[i].y, U[i].y);
    }
  }
  __syncthreads();
  float* ya = p.y1 + (size_t)env_a * npts;
  float* yb = p.y1 + (size_t)env_b * npts;
  for (int i = tid; i < npts; i += kGenThreads) {
    ya[i] = U[i].x;
    if (b_valid) yb[i] = U[i].y;
  }
}

}  // namespace pdeopt
