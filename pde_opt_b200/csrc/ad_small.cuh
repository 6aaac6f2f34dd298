// Fused advection-diffusion stepper (forward) for 32x32 and 64x64 grids: the structure of
// ad_generic.cuh (two environments per complex field, 3 forward + 1 inverse transform per step)
// with the two-stage register butterflies of sifs_small.cuh.  64x64 is the size of the reference's own
// advection-diffusion runs (notebooks/run_advection_diffusion.ipynb, the deleted AdvectionDiffusionEnv
// quoted in notebooks/test_pde_RL.ipynb:129).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ad_generic.cuh"
#include "sifs_small.cuh"

namespace pdeopt {

template <int N>
struct AdSmallSmem {
  static constexpr int P = N + 1;
  float2 U[N * P], Z[N * P], ACC[N * P];
  float mt[N * N], lt[N * N];  // filter multiplier and -L in position order (mt rebuilt when dt changes)
  float kxp[N], kyp[N];        // 2 pi k (Nyquist zeroed) in position order
  float2 tw[N];
  int p2f[N];
  float2 ax[N], ex[N], ay[N], ey[N];
};

template <int N>
__global__ void __launch_bounds__(kSmallThreads, 1) ad_small_fwd_kernel(const __grid_constant__ AdGenParams gp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Sm = AdSmallSmem<N>;
  Sm& S = *reinterpret_cast<Sm*>(smem_raw);
  constexpr int P = Sm::P, NPTS = N * N, NT = kSmallThreads, TL = N / 2 + 1;
  constexpr int R1 = lf_r1(N), R2 = lf_r2(N);
  const AdParams& p = gp.a;
  const int tid = threadIdx.x;
  const int env_a = 2 * blockIdx.x;
  const int env_b = (env_a + 1 < p.batch) ? env_a + 1 : env_a;
  const bool b_valid = env_a + 1 < p.batch;
  const float* tabA = p.tabA;
  const float* tabL = p.tabA + TL * TL;
  const float* kxs = tabL + TL * TL;
  const float* kys = kxs + N;

  for (int i = tid; i < N; i += NT) {
    float s, c;
    sincospif(-2.0f * float(i) / float(N), &s, &c);
    S.tw[i] = make_float2(c, s);
    const int f = line_pos_to_freq(N, i);
    S.p2f[i] = f;
    S.kxp[i] = kxs[f];
    S.kyp[i] = kys[f];
  }
  {
    const float* ya = p.y0 + (size_t)env_a * NPTS;
    const float* yb = p.y0 + (size_t)env_b * NPTS;
    for (int i = tid; i < NPTS; i += NT) S.U[(i / N) * P + (i % N)] = make_float2(ya[i], yb[i]);
  }
  __syncthreads();
  for (int i = tid; i < NPTS; i += NT) {
    const int kx = S.p2f[i / N], ky = S.p2f[i % N];
    const int fx = kx <= N / 2 ? kx : N - kx, fy = ky <= N / 2 ? ky : N - ky;
    S.lt[i] = -tabL[fx * TL + fy];
  }
  auto fwd2d = [&]() {
    ss_stage<N, N, P, true, R1, N, false>(S.Z, S.tw);
    ss_stage<N, N, P, true, R2, N / R1, false>(S.Z, S.tw);
    ss_stage<N, N, P, false, R1, N, false>(S.Z, S.tw);
    ss_stage<N, N, P, false, R2, N / R1, false>(S.Z, S.tw);
  };
  int cur_seg = -1;
  float dt_tab = __int_as_float(0x7fc00000);
  for (int k = 0; k < p.ksteps; ++k) {
    const int seg = ad_seg(p, k);
    const float dt = p.dt[k];
    if (seg != cur_seg || dt != dt_tab) {
      __syncthreads();
      if (seg != cur_seg) {
        const float* ca = p.ctrl + ((size_t)env_a * p.nseg + seg) * kAdCtrl;
        const float* cb = p.ctrl + ((size_t)env_b * p.nseg + seg) * kAdCtrl;
        for (int i = tid; i < 2 * N; i += NT) {
          const bool isx = i < N;
          const int q = isx ? i : i - N;
          const float pos = isx ? (p.lo_x + (q + 0.5f) * p.hx) : (p.lo_y + (q + 0.5f) * p.hy);
          const float2 d = make_float2(pos - (isx ? ca[0] : ca[1]), pos - (isx ? cb[0] : cb[1]));
          const float2 e = make_float2(expf(-d.x * d.x / (2.0f * ca[3])), expf(-d.y * d.y / (2.0f * cb[3])));
          const float2 a = make_float2(ca[2] * (-d.x / ca[3] * e.x), cb[2] * (-d.y / cb[3] * e.y));
          if (isx) { S.ax[q] = a; S.ex[q] = e; } else { S.ay[q] = a; S.ey[q] = e; }
        }
        cur_seg = seg;
      }
      if (dt != dt_tab) {
        for (int i = tid; i < NPTS; i += NT) {
          const int kx = S.p2f[i / N], ky = S.p2f[i % N];
          const int fx = kx <= N / 2 ? kx : N - kx, fy = ky <= N / 2 ? ky : N - ky;
          S.mt[i] = __fdividef(1.0f / float(NPTS), fmaf(dt, tabA[fx * TL + fy], 1.0f));
        }
        dt_tab = dt;
      }
      __syncthreads();
    }
    // ---- ACC = -L F[u] ----
    for (int i = tid; i < NPTS; i += NT) {
      const int a = (i / N) * P + (i % N);
      S.Z[a] = S.U[a];
    }
    __syncthreads();
    fwd2d();
    for (int i = tid; i < NPTS; i += NT) {
      const int a = (i / N) * P + (i % N);
      const float l = S.lt[i];
      S.ACC[a] = make_float2(S.Z[a].x * l, S.Z[a].y * l);
    }
    __syncthreads();
    // ---- ACC += -i kx F[vx u] ----
    for (int i = tid; i < NPTS; i += NT) {
      const int r = i / N, c = i % N, a = r * P + c;
      S.Z[a] = f2mul(f2mul(S.U[a], S.ax[r]), S.ey[c]);
    }
    __syncthreads();
    fwd2d();
    for (int i = tid; i < NPTS; i += NT) {
      const int a = (i / N) * P + (i % N);
      const float kk = S.kxp[i / N];
      S.ACC[a] = make_float2(fmaf(S.Z[a].y, kk, S.ACC[a].x), fmaf(-S.Z[a].x, kk, S.ACC[a].y));
    }
    __syncthreads();
    // ---- ACC += -i ky F[vy u];  filter;  inverse ----
    for (int i = tid; i < NPTS; i += NT) {
      const int r = i / N, c = i % N, a = r * P + c;
      S.Z[a] = f2mul(f2mul(S.U[a], S.ex[r]), S.ay[c]);
    }
    __syncthreads();
    fwd2d();
    for (int i = tid; i < NPTS; i += NT) {
      const int a = (i / N) * P + (i % N);
      const float kk = S.kyp[i % N], m = S.mt[i];
      S.Z[a] = make_float2(fmaf(S.Z[a].y, kk, S.ACC[a].x) * m, fmaf(-S.Z[a].x, kk, S.ACC[a].y) * m);
    }
    __syncthreads();
    ss_stage<N, N, P, false, R2, N / R1, true>(S.Z, S.tw);
    ss_stage<N, N, P, false, R1, N, true>(S.Z, S.tw);
    ss_stage<N, N, P, true, R2, N / R1, true>(S.Z, S.tw);
    ss_stage<N, N, P, true, R1, N, true>(S.Z, S.tw);
    for (int i = tid; i < NPTS; i += NT) {
      const int a = (i / N) * P + (i % N);
      S.U[a].x = fmaf(dt, S.Z[a].x, S.U[a].x);
      S.U[a].y = fmaf(dt, S.Z[a].y, S.U[a].y);
    }
    __syncthreads();
  }
  float* ya = p.y1 + (size_t)env_a * NPTS;
  float* yb = p.y1 + (size_t)env_b * NPTS;
  for (int i = tid; i < NPTS; i += NT) {
    const float2 v = S.U[(i / N) * P + (i % N)];
    ya[i] = v.x;
    if (b_valid) yb[i] = v.y;
  }
}

}  // namespace pdeopt
