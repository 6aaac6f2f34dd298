// Generic-size fused SIFS stepper: any power-of-two nx, ny with nx*ny <= 8192 (64x64, 64x128,
// 256x1, ...).  Same arithmetic as sifs128.cuh — pair trick z = u_a + i u_b, forward DIF /
// inverse DIT, multiplier applied in bit-reversed position — but written for generality, not
// speed: radix-2 stages in shared memory with one __syncthreads per stage.  It serves the sizes
// the tuned 128x128 kernel does not (BASELINE config 1: Allen-Cahn 64x64; the reference's own
// 256x1 known-answer test, tests/test_solvers.py:21-61) and is an independent cross-check of it.
//
// Replaces solvers.py:56-70 + cahn_hilliard.py:89-109 / allen_cahn.py:81-84 (see sifs128.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pointwise.cuh"
#include "sifs128.cuh"

namespace pdeopt {

constexpr int kGenThreads = 256;
constexpr int kGenMaxPts = 8192;

struct GenParams {
  SifsParams s;        // shared fields (pointers, dt[], pointwise params, geometry)
  int nx, ny, lognx, logny;
};

__device__ __forceinline__ int brev_rt(int v, int bits) { return (int)(__brev((unsigned)v) >> (32 - bits)) & ((1 << bits) - 1); }

// One radix-2 stage over axis with `len` = 1 << loglen elements and element stride `stride`
// (other-axis size = other, other-axis stride = ostride).  DIF (forward) or DIT (inverse).
template <bool INV>
__device__ __forceinline__ void r2_stage(float2* Z, const float2* tw, int twstep_shift, int half, int loglen, int stride,
                                         int other, int ostride, int npts) {
  // butterflies: for block base b (multiple of 2*half), j in [0, half): (b + j, b + j + half).
  // `other` (the size of the other axis) and `half` are powers of two: shifts and masks, no divisions.
  const int len = 1 << loglen;
  const int nb = npts / 2;
  const int oshift = 31 - __clz(other), omask = other - 1;
  for (int i = threadIdx.x; i < nb; i += kGenThreads) {
    const int o = i & omask;          // position along the other axis
    const int k = i >> oshift;        // butterfly index along this axis, 0 .. len/2-1
    const int j = k & (half - 1);
    const int base = ((k - j) << 1) + j;
    const int ia = base * stride + o * ostride, ib = ia + half * stride;
    const float2 a = Z[ia], b = Z[ib];
    // twiddle w_{2*half}^j = w_len^(j * len/(2*half))
    float2 w = tw[(j << twstep_shift) & (len - 1)];
    if (INV) w.y = -w.y;
    if (!INV) {
      Z[ia] = cadd(a, b);
      Z[ib] = cmul(csub(a, b), w);
    } else {
      const float2 t = cmul(b, w);
      Z[ia] = cadd(a, t);
      Z[ib] = csub(a, t);
    }
  }
}

template <int EQ>
__global__ void __launch_bounds__(kGenThreads, 1) sifs_generic_kernel(const __grid_constant__ GenParams gp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const SifsParams& p = gp.s;
  const int nx = gp.nx, ny = gp.ny, npts = nx * ny;
  float2* U = reinterpret_cast<float2*>(smem_raw);  // state pair
  float2* Z = U + npts;                             // work / mu buffer
  float2* twx = Z + npts;                           // w_nx^e, e < nx
  float2* twy = twx + nx;                           // w_ny^e
  float2* gx = twy + ny;
  float2* gy = gx + nx;
  float2* red = gy + ny;
  const int tid = threadIdx.x;
  const int env_a = 2 * blockIdx.x;
  const int env_b = (env_a + 1 < p.batch) ? env_a + 1 : env_a;
  const bool b_valid = env_a + 1 < p.batch;

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
  float2 w_off = make_float2(0.f, 0.f);
  const bool has_bump = p.ctrl != nullptr;
  if (has_bump) {
    const float* ca = p.ctrl + (size_t)env_a * kNCtrl;
    const float* cb = p.ctrl + (size_t)env_b * kNCtrl;
    w_off = make_float2(ca[0], cb[0]);
    for (int i = tid; i < nx + ny; i += kGenThreads) {
      const bool isx = i < nx;
      const int q = isx ? i : i - nx;
      const float pos = isx ? (p.lo_x + (q + 0.5f) * p.hx) : (p.lo_y + (q + 0.5f) * p.hy);
      const float da = pos - (isx ? ca[2] : ca[3]), db = pos - (isx ? cb[2] : cb[3]);
      const float ia = 0.5f / (ca[4] * ca[4]), ib = 0.5f / (cb[4] * cb[4]);
      float2 out;
      out.x = (ca[1] != 0.f ? expf(-da * da * ia) : 0.f) * (isx ? ca[1] : 1.0f);
      out.y = (cb[1] != 0.f ? expf(-db * db * ib) : 0.f) * (isx ? cb[1] : 1.0f);
      if (isx) gx[q] = out; else gy[q] = out;
    }
  }
  {
    const float* ya = p.y0 + (size_t)env_a * npts;
    const float* yb = p.y0 + (size_t)env_b * npts;
    for (int i = tid; i < npts; i += kGenThreads) U[i] = make_float2(ya[i], yb[i]);
  }
  __syncthreads();

  const int nsteps = (p.mode == MODE_RHS_ONLY) ? 1 : p.ksteps;
  for (int k = 0; k < nsteps; ++k) {
    // ---- RHS: mu and D into Z (mu) ... two-stage stencil through shared memory ----
    if (p.mode == MODE_GIVEN_F) {
      const float* fa = p.f0 + (size_t)env_a * npts;
      const float* fb = p.f0 + (size_t)env_b * npts;
      for (int i = tid; i < npts; i += kGenThreads) Z[i] = make_float2(fa[i], fb[i]);
      __syncthreads();
    } else {
      for (int i = tid; i < npts; i += kGenThreads) {
        const int r = i >> gp.logny, c = i & (ny - 1);
        const int rp = (r + 1) & (nx - 1), rm = (r + nx - 1) & (nx - 1), cp = (c + 1) & (ny - 1), cm = (c + ny - 1) & (ny - 1);
        const float2 u0 = U[i], up = U[rp * ny + c], um = U[rm * ny + c], ur = U[r * ny + cp], ul = U[r * ny + cm];
        float2 lap;
        lap.x = ((up.x - 2.0f * u0.x) + um.x) * p.inv_hx2 + ((ur.x - 2.0f * u0.x) + ul.x) * p.inv_hy2;
        lap.y = ((up.y - 2.0f * u0.y) + um.y) * p.inv_hx2 + ((ur.y - 2.0f * u0.y) + ul.y) * p.inv_hy2;
        float ma = mu_h<MU_RUNTIME>(u0.x, p.pw, w_off.x), mb = mu_h<MU_RUNTIME>(u0.y, p.pw, w_off.y);
        if (has_bump) {
          ma = fmaf(gx[r].x, gy[c].x, ma);
          mb = fmaf(gx[r].y, gy[c].y, mb);
        }
        Z[i] = make_float2(ma - p.kappa * lap.x, mb - p.kappa * lap.y);
      }
      __syncthreads();
      float2 f[kGenMaxPts / kGenThreads];
      int n = 0;
      for (int i = tid; i < npts; i += kGenThreads, ++n) {
        const int r = i >> gp.logny, c = i & (ny - 1);
        const float2 u0 = U[i], m0 = Z[i];
        const float2 D0 = make_float2(mob<MOB_RUNTIME>(u0.x, p.pw), mob<MOB_RUNTIME>(u0.y, p.pw));
        if (EQ == EQ_AC) {
          f[n] = make_float2(-D0.x * m0.x, -D0.y * m0.y);  // allen_cahn.py:84
        } else {
          const int rp = (r + 1) & (nx - 1), rm = (r + nx - 1) & (nx - 1), cp = (c + 1) & (ny - 1), cm = (c + ny - 1) & (ny - 1);
          const int ixp = rp * ny + c, ixm = rm * ny + c, iyp = r * ny + cp, iym = r * ny + cm;
          const float2 uxp = U[ixp], uxm = U[ixm], uyp = U[iyp], uym = U[iym];
          const float2 mxp = Z[ixp], mxm = Z[ixm], myp = Z[iyp], mym = Z[iym];
          float2 out;
#define PDEOPT_GEN_FLUX(comp)                                                                                      \
  {                                                                                                                \
    const float Dxp = mob<MOB_RUNTIME>(uxp.comp, p.pw), Dxm = mob<MOB_RUNTIME>(uxm.comp, p.pw);                    \
    const float Dyp = mob<MOB_RUNTIME>(uyp.comp, p.pw), Dym = mob<MOB_RUNTIME>(uym.comp, p.pw);                    \
    const float Fx1 = (0.5f * (D0.comp + Dxp)) * ((mxp.comp - m0.comp) * p.inv_hx);                                \
    const float Fx0 = (0.5f * (Dxm + D0.comp)) * ((m0.comp - mxm.comp) * p.inv_hx);                                \
    const float Fy1 = (0.5f * (D0.comp + Dyp)) * ((myp.comp - m0.comp) * p.inv_hy);                                \
    const float Fy0 = (0.5f * (Dym + D0.comp)) * ((m0.comp - mym.comp) * p.inv_hy);                                \
    out.comp = (Fx1 - Fx0) * p.inv_hx + (Fy1 - Fy0) * p.inv_hy;                                                    \
  }
          PDEOPT_GEN_FLUX(x)
          PDEOPT_GEN_FLUX(y)
#undef PDEOPT_GEN_FLUX
          f[n] = out;
        }
      }
      __syncthreads();
      n = 0;
      for (int i = tid; i < npts; i += kGenThreads, ++n) Z[i] = f[n];
      __syncthreads();
    }
    if (p.mode == MODE_RHS_ONLY) break;

    // ---- forward 2-D FFT (DIF, natural in, bit-reversed out per axis) ----
    for (int s = 0; s < gp.logny; ++s) {  // axis 1 (stride 1)
      r2_stage<false>(Z, twy, s, ny >> (s + 1), gp.logny, 1, nx, ny, npts);
      __syncthreads();
    }
    for (int s = 0; s < gp.lognx; ++s) {  // axis 0 (stride ny)
      r2_stage<false>(Z, twx, s, nx >> (s + 1), gp.lognx, ny, ny, 1, npts);
      __syncthreads();
    }
    // ---- multiplier at bit-reversed positions ----
    const float dt = p.dt[k];
    const float inv_n = 1.0f / float(npts);
    for (int i = tid; i < npts; i += kGenThreads) {
      const int r = i >> gp.logny, c = i & (ny - 1);
      const int kx = brev_rt(r, gp.lognx), ky = brev_rt(c, gp.logny);
      const int fx = kx <= nx / 2 ? kx : nx - kx, fy = ky <= ny / 2 ? ky : ny - ky;
      const float m = __fdividef(inv_n, fmaf(dt, p.symbol[fx * (ny / 2 + 1) + fy], 1.0f));
      Z[i] = make_float2(Z[i].x * m, Z[i].y * m);
    }
    __syncthreads();
    // ---- inverse (DIT, bit-reversed in, natural out) ----
    for (int s = gp.lognx - 1; s >= 0; --s) {
      r2_stage<true>(Z, twx, s, nx >> (s + 1), gp.lognx, ny, ny, 1, npts);
      __syncthreads();
    }
    for (int s = gp.logny - 1; s >= 0; --s) {
      r2_stage<true>(Z, twy, s, ny >> (s + 1), gp.logny, 1, nx, ny, npts);
      __syncthreads();
    }
    for (int i = tid; i < npts; i += kGenThreads) {
      U[i].x = fmaf(dt, Z[i].x, U[i].x);  // solvers.py:63
      U[i].y = fmaf(dt, Z[i].y, U[i].y);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const float2* src = (p.mode == MODE_RHS_ONLY) ? Z : U;
  float* ya = p.y1 + (size_t)env_a * npts;
  float* yb = p.y1 + (size_t)env_b * npts;
  float2 sum = make_float2(0.f, 0.f);
  for (int i = tid; i < npts; i += kGenThreads) {
    const float2 v = src[i];
    ya[i] = v.x;
    if (b_valid) yb[i] = v.y;
    sum.x += v.x;
    sum.y += v.y;
    if (p.obs != nullptr) {
      p.obs[(size_t)env_a * npts + i] = (uint8_t)rintf(__saturatef((v.x - p.obs_lo) * p.obs_scale) * 255.0f);
      if (b_valid) p.obs[(size_t)env_b * npts + i] = (uint8_t)rintf(__saturatef((v.y - p.obs_lo) * p.obs_scale) * 255.0f);
    }
  }
  if (p.reward != nullptr) {
    auto block_sum = [&](float2 v) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
      }
      __syncthreads();
      if ((tid & 31) == 0) red[tid >> 5] = v;
      __syncthreads();
      float2 t = make_float2(0.f, 0.f);
      for (int w = 0; w < kGenThreads / 32; ++w) {
        t.x += red[w].x;
        t.y += red[w].y;
      }
      return t;
    };
    const float inv_n = 1.0f / float(npts);
    const float2 tot = block_sum(sum);
    const float2 mean = make_float2(tot.x * inv_n, tot.y * inv_n);
    float2 sq = make_float2(0.f, 0.f);
    for (int i = tid; i < npts; i += kGenThreads) {
      const float da = src[i].x - mean.x, db = src[i].y - mean.y;
      sq.x = fmaf(da, da, sq.x);
      sq.y = fmaf(db, db, sq.y);
    }
    const float2 tsq = block_sum(sq);
    if (tid == 0) {
      p.reward[2 * env_a] = mean.x;
      p.reward[2 * env_a + 1] = tsq.x * inv_n;
      if (b_valid) {
        p.reward[2 * env_b] = mean.y;
        p.reward[2 * env_b + 1] = tsq.y * inv_n;
      }
    }
  }
}

inline size_t gen_smem_bytes(int nx, int ny) {
  return sizeof(float2) * (size_t)(2 * nx * ny + 2 * nx + 2 * ny + kGenThreads / 32 + 4);
}

}  // namespace pdeopt
