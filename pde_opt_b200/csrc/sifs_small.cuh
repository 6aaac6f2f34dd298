// Fused K-step semi-implicit stepper for small square grids (32x32, 64x64): one CTA of 512 threads
// per pair of environments, the field resident in shared memory, radix-8 / radix-4 register
// butterflies (two stages per axis) instead of the twelve radix-2 stages of sifs_generic.cuh.
// These are the sizes the reference's own notebooks run (32x32 / 64x64 Cahn-Hilliard and
// advection-diffusion; BASELINE config 1 is Allen-Cahn 64x64, single environment, 1000 steps), where
// a single environment is latency-bound: the kernel is built for few barriers per step.
//
// Replaces SemiImplicitFourierSpectral.step (pde_opt/numerics/solvers.py:56-70) with
// CahnHilliard2DPeriodic.rhs_fd (cahn_hilliard.py:89-109) / AllenCahn2DPeriodic.rhs_fd
// (allen_cahn.py:81-84); same arithmetic and expression order as sifs_generic.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "linefft.cuh"
#include "sifs_generic.cuh"

namespace pdeopt {

constexpr int kSmallThreads = 512;

// One radix-R stage with span S over the lines of a 2-D shared-memory array with row pitch PITCH:
// ALONG_Y: the transform runs along y (axis 1), lines are the NL rows; otherwise along x, lines are
// the NL columns.  Lanes run across lines: pitch-strided (odd pitch) or unit-strided, conflict free.
template <int N, int NL, int PITCH, bool ALONG_Y, int R, int S, bool INV>
__device__ __forceinline__ void ss_stage(float2* __restrict__ Z, const float2* __restrict__ tw) {
  if constexpr (R > 1) {
    constexpr int SUB = S / R, NB = N / R, LOG2R = ilog2(R);
    for (int w = threadIdx.x; w < NB * NL; w += kSmallThreads) {
      const int line = w % NL, u = w / NL;
      const int j = u % SUB, base = (u / SUB) * S + j;
      auto at = [&](int pos) -> float2& { return ALONG_Y ? Z[line * PITCH + pos] : Z[pos * PITCH + line]; };
      float2 x[R];
      if constexpr (!INV) {
#pragma unroll
        for (int m = 0; m < R; ++m) x[m] = at(base + m * SUB);
        Dif<R, 1, false>::run(x);
        static_for<0, R>([&](auto pc) {
          constexpr int p = decltype(pc)::value;
          constexpr int k = brev<LOG2R>(p);
          float2 v = x[p];
          if constexpr (k != 0 && SUB > 1) v = cmul(v, tw[(j * k * (N / S)) & (N - 1)]);
          at(base + k * SUB) = v;
        });
      } else {
        static_for<0, R>([&](auto pc) {
          constexpr int p = decltype(pc)::value;
          constexpr int k = brev<LOG2R>(p);
          float2 v = at(base + k * SUB);
          if constexpr (k != 0 && SUB > 1) v = cmulc(v, tw[(j * k * (N / S)) & (N - 1)]);
          x[p] = v;
        });
        Dit<R, 1, true>::run(x);
#pragma unroll
        for (int m = 0; m < R; ++m) at(base + m * SUB) = x[m];
      }
    }
    __syncthreads();
  }
}

template <int N>
struct SmallSmem {
  static constexpr int P = N + 1;
  float2 U[N * P], Z[N * P], M[N * P], Dm[N * P];
  float mt[N * N];   // multiplier in position order on both axes, rebuilt when dt changes
  float2 tw[N];
  int p2f[N];
  float2 gx[N], gy[N];
  float2 red[kSmallThreads / 32];
};

template <int N, int EQ>
__global__ void __launch_bounds__(kSmallThreads, 1) sifs_small_kernel(const __grid_constant__ GenParams gp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Sm = SmallSmem<N>;
  Sm& S = *reinterpret_cast<Sm*>(smem_raw);
  constexpr int P = Sm::P, NPTS = N * N, NT = kSmallThreads;
  constexpr int R1 = lf_r1(N), R2 = lf_r2(N);
  static_assert(lf_r3(N) == 1, "two-stage plans only");
  const SifsParams& p = gp.s;
  const int tid = threadIdx.x;
  const int env_a = 2 * blockIdx.x;
  const int env_b = (env_a + 1 < p.batch) ? env_a + 1 : env_a;
  const bool b_valid = env_a + 1 < p.batch;

  for (int i = tid; i < N; i += NT) {
    float s, c;
    sincospif(-2.0f * float(i) / float(N), &s, &c);
    S.tw[i] = make_float2(c, s);
    S.p2f[i] = line_pos_to_freq(N, i);
  }
  float2 w_off = make_float2(0.f, 0.f);
  const bool has_bump = p.ctrl != nullptr;
  if (has_bump) {
    const float* ca = p.ctrl + (size_t)env_a * kNCtrl;
    const float* cb = p.ctrl + (size_t)env_b * kNCtrl;
    w_off = make_float2(ca[0], cb[0]);
    for (int i = tid; i < 2 * N; i += NT) {
      const bool isx = i < N;
      const int q = isx ? i : i - N;
      const float pos = isx ? (p.lo_x + (q + 0.5f) * p.hx) : (p.lo_y + (q + 0.5f) * p.hy);
      const float da = pos - (isx ? ca[2] : ca[3]), db = pos - (isx ? cb[2] : cb[3]);
      const float ia = 0.5f / (ca[4] * ca[4]), ib = 0.5f / (cb[4] * cb[4]);
      float2 out;
      out.x = (ca[1] != 0.f ? expf(-da * da * ia) : 0.f) * (isx ? ca[1] : 1.0f);
      out.y = (cb[1] != 0.f ? expf(-db * db * ib) : 0.f) * (isx ? cb[1] : 1.0f);
      if (isx) S.gx[q] = out; else S.gy[q] = out;
    }
  }
  {
    const float* ya = p.y0 + (size_t)env_a * NPTS;
    const float* yb = p.y0 + (size_t)env_b * NPTS;
    for (int i = tid; i < NPTS; i += NT) S.U[(i / N) * P + (i % N)] = make_float2(ya[i], yb[i]);
  }
  __syncthreads();

  float dt_tab = __int_as_float(0x7fc00000);
  const int nsteps = (p.mode == MODE_RHS_ONLY) ? 1 : p.ksteps;
  for (int k = 0; k < nsteps; ++k) {
    if (p.mode != MODE_RHS_ONLY && p.dt[k] != dt_tab) {
      // 1 / (N^2 (1 + dt A sigma)) in position order (solvers.py:62-63); the barriers of the RHS phase
      // below order these writes before the multiply
      dt_tab = p.dt[k];
      for (int i = tid; i < NPTS; i += NT) {
        const int kx = S.p2f[i / N], ky = S.p2f[i % N];
        const int fx = kx <= N / 2 ? kx : N - kx, fy = ky <= N / 2 ? ky : N - ky;
        S.mt[i] = __fdividef(1.0f / float(NPTS), fmaf(dt_tab, p.symbol[fx * (N / 2 + 1) + fy], 1.0f));
      }
    }
    // ---- RHS into Z ----
    if (p.mode == MODE_GIVEN_F) {
      const float* fa = p.f0 + (size_t)env_a * NPTS;
      const float* fb = p.f0 + (size_t)env_b * NPTS;
      for (int i = tid; i < NPTS; i += NT) S.Z[(i / N) * P + (i % N)] = make_float2(fa[i], fb[i]);
      __syncthreads();
    } else {
      for (int i = tid; i < NPTS; i += NT) {
        const int r = i / N, c = i % N;
        const int rp = (r + 1) & (N - 1), rm = (r + N - 1) & (N - 1), cp = (c + 1) & (N - 1), cm = (c + N - 1) & (N - 1);
        const float2 u0 = S.U[r * P + c], up = S.U[rp * P + c], um = S.U[rm * P + c], ur = S.U[r * P + cp], ul = S.U[r * P + cm];
        float2 lap;
        lap.x = ((up.x - 2.0f * u0.x) + um.x) * p.inv_hx2 + ((ur.x - 2.0f * u0.x) + ul.x) * p.inv_hy2;
        lap.y = ((up.y - 2.0f * u0.y) + um.y) * p.inv_hx2 + ((ur.y - 2.0f * u0.y) + ul.y) * p.inv_hy2;
        float ma = mu_h<MU_RUNTIME>(u0.x, p.pw, w_off.x), mb = mu_h<MU_RUNTIME>(u0.y, p.pw, w_off.y);
        if (has_bump) {
          ma = fmaf(S.gx[r].x, S.gy[c].x, ma);
          mb = fmaf(S.gx[r].y, S.gy[c].y, mb);
        }
        const float2 mu = make_float2(ma - p.kappa * lap.x, mb - p.kappa * lap.y);
        const float2 D0 = make_float2(mob<MOB_RUNTIME>(u0.x, p.pw), mob<MOB_RUNTIME>(u0.y, p.pw));
        if (EQ == EQ_AC) {
          S.Z[r * P + c] = make_float2(-D0.x * mu.x, -D0.y * mu.y);  // allen_cahn.py:84
        } else {
          S.M[r * P + c] = mu;
          S.Dm[r * P + c] = D0;
        }
      }
      __syncthreads();
      if (EQ == EQ_CH) {
        for (int i = tid; i < NPTS; i += NT) {
          const int r = i / N, c = i % N;
          const int rp = (r + 1) & (N - 1), rm = (r + N - 1) & (N - 1), cp = (c + 1) & (N - 1), cm = (c + N - 1) & (N - 1);
          const float2 m0 = S.M[r * P + c], D0 = S.Dm[r * P + c];
          const float2 mxp = S.M[rp * P + c], mxm = S.M[rm * P + c], myp = S.M[r * P + cp], mym = S.M[r * P + cm];
          const float2 Dxp = S.Dm[rp * P + c], Dxm = S.Dm[rm * P + c], Dyp = S.Dm[r * P + cp], Dym = S.Dm[r * P + cm];
          float2 out;
#define PDEOPT_SMALL_FLUX(comp)                                                              \
  {                                                                                          \
    const float Fx1 = (0.5f * (D0.comp + Dxp.comp)) * ((mxp.comp - m0.comp) * p.inv_hx);      \
    const float Fx0 = (0.5f * (Dxm.comp + D0.comp)) * ((m0.comp - mxm.comp) * p.inv_hx);      \
    const float Fy1 = (0.5f * (D0.comp + Dyp.comp)) * ((myp.comp - m0.comp) * p.inv_hy);      \
    const float Fy0 = (0.5f * (Dym.comp + D0.comp)) * ((m0.comp - mym.comp) * p.inv_hy);      \
    out.comp = (Fx1 - Fx0) * p.inv_hx + (Fy1 - Fy0) * p.inv_hy;                               \
  }
          PDEOPT_SMALL_FLUX(x)
          PDEOPT_SMALL_FLUX(y)
#undef PDEOPT_SMALL_FLUX
          S.Z[r * P + c] = out;
        }
        __syncthreads();
      }
    }
    if (p.mode == MODE_RHS_ONLY) break;
    // ---- forward: along y then along x (DIF, position order out) ----
    ss_stage<N, N, P, true, R1, N, false>(S.Z, S.tw);
    ss_stage<N, N, P, true, R2, N / R1, false>(S.Z, S.tw);
    ss_stage<N, N, P, false, R1, N, false>(S.Z, S.tw);
    ss_stage<N, N, P, false, R2, N / R1, false>(S.Z, S.tw);
    const float dt = p.dt[k];
    for (int i = tid; i < NPTS; i += NT) {
      const int a = (i / N) * P + (i % N);
      const float m = S.mt[i];
      S.Z[a] = make_float2(S.Z[a].x * m, S.Z[a].y * m);
    }
    __syncthreads();
    // ---- inverse: along x then along y (DIT, natural order out) ----
    ss_stage<N, N, P, false, R2, N / R1, true>(S.Z, S.tw);
    ss_stage<N, N, P, false, R1, N, true>(S.Z, S.tw);
    ss_stage<N, N, P, true, R2, N / R1, true>(S.Z, S.tw);
    ss_stage<N, N, P, true, R1, N, true>(S.Z, S.tw);
    for (int i = tid; i < NPTS; i += NT) {
      const int a = (i / N) * P + (i % N);
      S.U[a].x = fmaf(dt, S.Z[a].x, S.U[a].x);  // solvers.py:63
      S.U[a].y = fmaf(dt, S.Z[a].y, S.U[a].y);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const float2* src = (p.mode == MODE_RHS_ONLY) ? S.Z : S.U;
  float* ya = p.y1 + (size_t)env_a * NPTS;
  float* yb = p.y1 + (size_t)env_b * NPTS;
  float2 sum = make_float2(0.f, 0.f);
  for (int i = tid; i < NPTS; i += NT) {
    const float2 v = src[(i / N) * P + (i % N)];
    ya[i] = v.x;
    if (b_valid) yb[i] = v.y;
    sum.x += v.x;
    sum.y += v.y;
    if (p.obs != nullptr) {
      p.obs[(size_t)env_a * NPTS + i] = (uint8_t)rintf(__saturatef((v.x - p.obs_lo) * p.obs_scale) * 255.0f);
      if (b_valid) p.obs[(size_t)env_b * NPTS + i] = (uint8_t)rintf(__saturatef((v.y - p.obs_lo) * p.obs_scale) * 255.0f);
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
      if ((tid & 31) == 0) S.red[tid >> 5] = v;
      __syncthreads();
      float2 t = make_float2(0.f, 0.f);
      for (int w = 0; w < NT / 32; ++w) {
        t.x += S.red[w].x;
        t.y += S.red[w].y;
      }
      return t;
    };
    const float inv_n = 1.0f / float(NPTS);
    const float2 tot = block_sum(sum);
    const float2 mean = make_float2(tot.x * inv_n, tot.y * inv_n);
    float2 sq = make_float2(0.f, 0.f);
    for (int i = tid; i < NPTS; i += NT) {
      const float2 v = src[(i / N) * P + (i % N)];
      const float da = v.x - mean.x, db = v.y - mean.y;
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

}  // namespace pdeopt
