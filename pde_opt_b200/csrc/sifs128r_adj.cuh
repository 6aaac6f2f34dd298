// Fused K-step discrete adjoint of the semi-implicit finite-difference phase-field step on 128x128 grids:
// ONE environment per 256-thread CTA, the cotangent stays on chip for all K steps of the launch, the
// coefficient cotangents are reduced in the CTA and written once per environment (sm_100a).
//
// This is the backward half of the differentiable rollout (`sifs_rollout_bwd` of SURVEY 8b): what reverse-mode
// differentiation through diffeqsolve (diffrax RecursiveCheckpointAdjoint, pde_model.py:226-323) yields for
// PDEModel.mse when the optimised leaves are the closure coefficients.  Per step, in reverse order
// (ch_adjoint.cuh has the derivation and the streaming reference implementation this kernel is tested against):
//     w      = dt G lam1                       filter of the forward step (solvers.py:62-63), rfft128.cuh
//     CH:  mu_bar = div( D_face grad_face(w) ),  D_bar = -1/2 sum_faces grad_face(w) grad_face(mu)
//     AC:  mu_bar = -R(u) w,                     D_bar = -w mu
//     lam0   = lam1 + mu_h'(u) mu_bar - kappa lap(mu_bar) + D'(u) D_bar
//     g_mu  += sum d mu_h / d theta (u) mu_bar,   g_D += sum d D / d theta (u) D_bar
// with u the state at the START of the step, read from the trajectory pdeopt_sifs_rollout_fwd kept in HBM
// (64 KB per environment and step: the only HBM traffic of a step).
//
// On chip: lam parked in 128 TMEM columns (pass-A register arrangement, bit-reversed register order so that it
// feeds passA_fwd directly), W = 64 KB field buffer (exchange space of the transform, then w in the natural
// layout, then the stencil result), M = 64 KB (mu_bar between the two stencil sweeps), the 33 KB filter table.
// Sweep A: warp = 16 rows, lane = 4 columns, rolling three-row windows of u, mu, D, w down the rows (column
// neighbours by shuffle, u rows straight from global memory, w rows from W); sweep B: lap(mu_bar) from M.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ch_adjoint.cuh"  // mu_h_prime, mob_prime, pointwise families
#include "sifs128r.cuh"

namespace pdeopt {
namespace rf {

struct AdjParams {
  const float* traj;  // [ksteps][batch][128][128] state at the start of every step of this launch
  const float* lam1;  // [batch][128][128] cotangent after the last step of the launch
  float* lam0;        // [batch][128][128] cotangent before its first step (may alias lam1)
  double* gmu;        // [batch][16] accumulated (+=)
  double* gmob;       // [batch][16] accumulated (+=)
  const float* symbol;
  int batch, ksteps, eq;  // eq: EQ_CH / EQ_AC
  float inv_hx, inv_hy, inv_hx2, inv_hy2, kappa;
  PointwiseParams pw;
  float dt[kMaxK];
};

struct __align__(1024) ASmem {
  float2 W[kRows * kH];
  float2 M[kRows * kH];
  float4 T[kTRows * kTCols];
  float2 twb[8 * 16];
  float2 tw64[32];
  float2 sc[32];
  double gacc[kThreadsR / 32][32];
  uint32_t tmem_base;
};

__device__ __forceinline__ void load_row4(uint32_t base, int r, int lane, float (&v)[4]) {
  const float4 a = ld4<0>(srow_addr(base, r & (kRows - 1), lane));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void store_row4(uint32_t base, int r, int lane, const float (&v)[4]) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow_addr(base, r & (kRows - 1), lane)), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3])
               : "memory");
}
__device__ __forceinline__ void load_grow4(const float* __restrict__ f, int r, int lane, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(f + (r & (kRows - 1)) * kCols + 4 * lane));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
// left / right column neighbours of the four values a lane holds
__device__ __forceinline__ void col_nbrs(const float (&a)[4], int lm1, int lp1, float (&L)[4], float (&R)[4]) {
  const float aL = shf(a[3], lm1), aR = shf(a[0], lp1);
  L[0] = aL; L[1] = a[0]; L[2] = a[1]; L[3] = a[2];
  R[0] = a[1]; R[1] = a[2]; R[2] = a[3]; R[3] = aR;
}

// mu_h(c), D(c) and their derivatives in one pass over the Legendre basis (one recurrence serves all four)
// NC: compile-time bound on the number of Legendre coefficients (the unrolled recurrence stops there: the sweep is
// instruction-fetch bound when all 16 terms are unrolled for every point, profiles/ncu_full_r2_sifs128r_adj_512x256.txt)
template <int NC>
__device__ __forceinline__ void pw_eval(float c, const PointwiseParams& pw, float& mu, float& D, float& mup, float& Dp) {
  const bool mleg = pw.mu_family == MU_LEGENDRE || pw.mu_family == MU_LEGENDRE_LOGPRIOR;
  const bool dleg = pw.mob_family == MOB_LEGENDRE_EXP;
  float sm = 0.f, dsm = 0.f, sd = 0.f, dsd = 0.f;
  if (mleg || dleg) {
    const float x = 2.0f * c - 1.0f;
    const int nm = mleg ? pw.mu_ncoef : 0, nd = dleg ? pw.mob_ncoef : 0;
    const int nmax = nm > nd ? nm : nd;
    float pp = 1.0f, pc = x, dpp = 0.0f, dpc = 1.0f;
    if (nm > 0) sm = pw.mu_coef[0];
    if (nd > 0) sd = pw.mob_coef[0];
    if (nm > 1) { sm = fmaf(pw.mu_coef[1], x, sm); dsm = pw.mu_coef[1]; }
    if (nd > 1) { sd = fmaf(pw.mob_coef[1], x, sd); dsd = pw.mob_coef[1]; }
    static_for<2, (NC > 2 ? NC : 2)>([&](auto nc) {
      constexpr int n = decltype(nc)::value;
      if (n < nmax) {
        const float pn = LegC<n>::a * x * pc - LegC<n>::b * pp;
        const float dpn = fmaf(float(2 * n - 1), pc, dpp);  // P_n' = P_{n-2}' + (2n-1) P_{n-1}
        if (n < nm) { sm = fmaf(pw.mu_coef[n], pn, sm); dsm = fmaf(pw.mu_coef[n], dpn, dsm); }
        if (n < nd) { sd = fmaf(pw.mob_coef[n], pn, sd); dsd = fmaf(pw.mob_coef[n], dpn, dsd); }
        pp = pc; pc = pn; dpp = dpc; dpc = dpn;
      }
    });
  }
  const float omc = 1.0f - c;
  switch (pw.mu_family) {
    case MU_DOUBLE_WELL: mu = c * c * c - c; mup = 3.0f * c * c - 1.0f; break;
    case MU_LOG:
      mu = __logf(__fdividef(c, omc)) + pw.mu_coef[0] * (1.0f - 2.0f * c);
      mup = __fdividef(1.0f, c * omc) - 2.0f * pw.mu_coef[0];
      break;
    case MU_LEGENDRE: mu = sm; mup = 2.0f * dsm; break;
    default: mu = sm + __logf(__fdividef(c, omc)); mup = 2.0f * dsm + __fdividef(1.0f, c * omc); break;
  }
  switch (pw.mob_family) {
    case MOB_CONST: D = pw.mob_coef[0]; Dp = 0.0f; break;
    case MOB_DEGENERATE: D = omc * c; Dp = 1.0f - 2.0f * c; break;
    case MOB_ONE_PLUS_SQ: D = 1.0f + c * c; Dp = 2.0f * c; break;
    default: D = __expf(sd); Dp = D * 2.0f * dsd; break;
  }
}

// d mu_h / d theta_n (c) * s accumulated into acc[0..15], d D / d theta_n (c) * t into acc[16..31]
template <int NC>
__device__ __forceinline__ void accumulate_coef(float (&acc)[32], float c, float Dval, float s, float t, const PointwiseParams& pw) {
  const bool mleg = pw.mu_family == MU_LEGENDRE || pw.mu_family == MU_LEGENDRE_LOGPRIOR;
  const bool dleg = pw.mob_family == MOB_LEGENDRE_EXP;
  if (pw.mu_family == MU_LOG) acc[0] = fmaf(1.0f - 2.0f * c, s, acc[0]);
  if (pw.mob_family == MOB_CONST) acc[16] += t;
  if (mleg || dleg) {
    const float x = 2.0f * c - 1.0f;
    const int nm = mleg ? pw.mu_ncoef : 0, nd = dleg ? pw.mob_ncoef : 0;
    const int nmax = nm > nd ? nm : nd;
    const float tD = Dval * t;
    float pp = 1.0f, pc = x;
    if (nm > 0) acc[0] += s;
    if (nd > 0) acc[16] += tD;
    if (nm > 1) acc[1] = fmaf(x, s, acc[1]);
    if (nd > 1) acc[17] = fmaf(x, tD, acc[17]);
    static_for<2, (NC > 2 ? NC : 2)>([&](auto nc) {
      constexpr int n = decltype(nc)::value;
      if (n < nmax) {
        const float pn = LegC<n>::a * x * pc - LegC<n>::b * pp;
        if (n < nm) acc[n] = fmaf(pn, s, acc[n]);
        if (n < nd) acc[16 + n] = fmaf(pn, tD, acc[16 + n]);
        pp = pc;
        pc = pn;
      }
    });
  }
}

template <int EQ, int NC>
__global__ void __launch_bounds__(kThreadsR, 1) sifs128r_adj_kernel(const __grid_constant__ AdjParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  ASmem& S = *reinterpret_cast<ASmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lm1 = (lane + 31) & 31, lp1 = (lane + 1) & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&S.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid < 128) {
    const int k1r = tid >> 3, n2r = tid & 7;
    float s, c;
    sincospif(-2.0f * float(n2r * k1r) / 128.0f, &s, &c);
    S.twb[tid] = make_float2(c, s);
  } else if (tid < 160) {
    float s, c;
    sincospif(-2.0f * float(tid - 128) / 64.0f, &s, &c);
    S.tw64[tid - 128] = make_float2(c, s);
  } else if (tid < 192) {
    float s, c;
    sincospif(2.0f * float(tid - 160) / 128.0f, &s, &c);
    S.sc[tid - 160] = make_float2(c, s);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // TMEM columns per warp quadrant: [0, 128) lam (64 per thread, two warps per lane quadrant), [128, 256) t1
  ParkR park;
  park.taddr = S.tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;

  const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(S.W);
  const uint32_t mbase = (uint32_t)__cvta_generic_to_shared(S.M);
  const RFft F(wbase, (uint32_t)__cvta_generic_to_shared(S.T), tid);
  float dt_tab = __int_as_float(0x7fc00000);
  const int r0 = warp * 16;
  const float hx2 = p.inv_hx2, hy2 = p.inv_hy2;

  for (int env = blockIdx.x; env < p.batch; env += gridDim.x) {
    S.gacc[warp][lane] = 0.0;
    // ---- prologue: lam1 -> natural layout -> pass-A registers (bit-reversed order) -> parked ----
    {
      const float* le = p.lam1 + (size_t)env * kRows * kCols;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v[4];
        load_grow4(le, r0 + i, lane, v);
        store_row4(wbase, r0 + i, lane, v);
      }
    }
    __syncthreads();
    float2 x[32];
    gather_nat<true>(F, x);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = x[ch * 8 + i];
      park.store(ch, v);
    }
    tmem_wait_st();
    __syncthreads();

    for (int k = p.ksteps - 1; k >= 0; --k) {
      const float dt = p.dt[k];
      const float* ue = p.traj + ((size_t)k * p.batch + env) * kRows * kCols;
      // the rows of u this warp will read in sweep A (r0-2 .. r0+17), pulled into L2 behind the transform
      if (lane < 20) asm volatile("prefetch.global.L2 [%0];" ::"l"(ue + ((r0 - 2 + lane) & (kRows - 1)) * kCols));
      if (lane < 20) asm volatile("prefetch.global.L2 [%0];" ::"l"(ue + ((r0 - 2 + lane) & (kRows - 1)) * kCols + 32));
      if (lane < 20) asm volatile("prefetch.global.L2 [%0];" ::"l"(ue + ((r0 - 2 + lane) & (kRows - 1)) * kCols + 64));
      if (lane < 20) asm volatile("prefetch.global.L2 [%0];" ::"l"(ue + ((r0 - 2 + lane) & (kRows - 1)) * kCols + 96));
      if (dt != dt_tab) {
        dt_tab = dt;
        build_table_r(S.T, p.symbol, S.sc, dt);
      }
      // ---- w = dt G lam1 (x holds lam1 in bit-reversed pass-A order) ----
      passA_fwd(F, x);
      __syncthreads();
      passB_fwd(F, S.twb, S.tw64, x);
      __syncthreads();
      passC_filter(F, x);
      __syncthreads();
      passB_inv(F, S.twb, S.tw64, x);
      __syncthreads();
      passA_inv(F, x);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = mul2(x[i], splat2(dt));
      __syncthreads();  // all exchange-layout reads are done before the natural layout is written
      scatter_nat(F, x);
      __syncthreads();

      // ---- sweep A: mu_bar -> M, t1 = mu_h'(u) mu_bar + D'(u) D_bar -> TMEM (16 values per group of 4 rows) ----
      float acc[32];
#pragma unroll
      for (int n = 0; n < 32; ++n) acc[n] = 0.f;
      {
        float um[4], u0[4], up[4];
        float mu_m[4], mu_0[4], mu_p[4], D_m[4], D_0[4], D_p[4], w_m[4], w_0[4], w_p[4], uc[4];
        float mq_0[4], mq_p[4], dq_0[4], dq_p[4];  // mu_h'(u), D'(u) of the window rows
        float t1g[4][4];
        load_grow4(ue, r0 - 2 + kRows, lane, um);
        load_grow4(ue, r0 - 1 + kRows, lane, u0);
        // one marching iteration: mu, D, w of row r0 + it enter the windows; if EMIT >= 0 the outputs of row
        // rho = r0 + it - 1 (windows (m, 0, p) = rows rho-1, rho, rho+1; uc = u of row rho) go to slot EMIT of t1g
        auto march = [&](int it, auto emit_c) {
          constexpr int EMIT = decltype(emit_c)::value;
          load_grow4(ue, r0 + it + 1 + kRows, lane, up);
          float uL[4], uR[4];
          col_nbrs(u0, lm1, lp1, uL, uR);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float lap = ((up[j] - 2.0f * u0[j]) + um[j]) * hx2 + ((uR[j] - 2.0f * u0[j]) + uL[j]) * hy2;
            float mh;
            pw_eval<NC>(u0[j], p.pw, mh, D_p[j], mq_p[j], dq_p[j]);
            mu_p[j] = mh - p.kappa * lap;
          }
          load_row4(wbase, r0 + it + kRows, lane, w_p);
          if constexpr (EMIT >= 0) {
            float mb[4], db[4];
            if (EQ == EQ_AC) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                mb[j] = -D_0[j] * w_0[j];
                db[j] = -w_0[j] * mu_0[j];
              }
            } else {
              float wL[4], wR[4], mL[4], mR[4], DL[4], DR[4];
              col_nbrs(w_0, lm1, lp1, wL, wR);
              col_nbrs(mu_0, lm1, lp1, mL, mR);
              col_nbrs(D_0, lm1, lp1, DL, DR);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float gwxp = (w_p[j] - w_0[j]) * p.inv_hx, gwxm = (w_0[j] - w_m[j]) * p.inv_hx;
                const float gwyp = (wR[j] - w_0[j]) * p.inv_hy, gwym = (w_0[j] - wL[j]) * p.inv_hy;
                const float gmxp = (mu_p[j] - mu_0[j]) * p.inv_hx, gmxm = (mu_0[j] - mu_m[j]) * p.inv_hx;
                const float gmyp = (mR[j] - mu_0[j]) * p.inv_hy, gmym = (mu_0[j] - mL[j]) * p.inv_hy;
                const float Dxp = 0.5f * (D_0[j] + D_p[j]), Dxm = 0.5f * (D_m[j] + D_0[j]);
                const float Dyp = 0.5f * (D_0[j] + DR[j]), Dym = 0.5f * (DL[j] + D_0[j]);
                mb[j] = (Dxp * gwxp - Dxm * gwxm) * p.inv_hx + (Dyp * gwyp - Dym * gwym) * p.inv_hy;
                db[j] = -0.5f * ((gwxp * gmxp + gwxm * gmxm) + (gwyp * gmyp + gwym * gmym));
              }
            }
            store_row4(mbase, r0 + it - 1, lane, mb);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              t1g[EMIT][j] = mq_0[j] * mb[j] + dq_0[j] * db[j];
              accumulate_coef<NC>(acc, uc[j], D_0[j], mb[j], db[j], p.pw);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mu_m[j] = mu_0[j]; mu_0[j] = mu_p[j];
            D_m[j] = D_0[j];   D_0[j] = D_p[j];
            w_m[j] = w_0[j];   w_0[j] = w_p[j];
            mq_0[j] = mq_p[j]; dq_0[j] = dq_p[j];
            uc[j] = u0[j];
            um[j] = u0[j];     u0[j] = up[j];
          }
        };
        march(-1, std::integral_constant<int, -1>{});
        march(0, std::integral_constant<int, -1>{});
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          static_for<0, 4>([&](auto ic) { march(1 + 4 * g + decltype(ic)::value, ic); });
          float2 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = make_float2(t1g[i >> 1][(i & 1) * 2], t1g[i >> 1][(i & 1) * 2 + 1]);
          tmem_st16(park.taddr + 128 + g * 16, v);
        }
        tmem_wait_st();
      }
      // coefficient cotangents of this step: warp reduction, then float64 accumulation per warp
#pragma unroll
      for (int n = 0; n < 32; ++n) {
        const bool used = (n < 16) ? (n < p.pw.mu_ncoef || (n == 0 && p.pw.mu_family == MU_LOG))
                                   : (n - 16 < p.pw.mob_ncoef || (n == 16 && p.pw.mob_family == MOB_CONST));
        if (used) {
          float v = acc[n];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) S.gacc[warp][n] += (double)v;
        }
      }
      __syncthreads();  // mu_bar rows of the neighbouring warps are in M; nobody reads w any more

      // ---- sweep B: s = t1 - kappa lap(mu_bar) -> W (natural layout) ----
      {
        float bm[4], b0[4], bp[4];
        load_row4(mbase, r0 - 1 + kRows, lane, bm);
        load_row4(mbase, r0, lane, b0);
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          float2 t1v[8];
          tmem_ld16(park.taddr + 128 + g * 16, t1v);
          static_for<0, 4>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            const int row = r0 + 4 * g + i;
            load_row4(mbase, row + 1, lane, bp);
            float bL[4], bR[4], sv[4];
            col_nbrs(b0, lm1, lp1, bL, bR);
            const float t1r[4] = {t1v[2 * i].x, t1v[2 * i].y, t1v[2 * i + 1].x, t1v[2 * i + 1].y};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float lap = ((bp[j] - 2.0f * b0[j]) + bm[j]) * hx2 + ((bR[j] - 2.0f * b0[j]) + bL[j]) * hy2;
              sv[j] = t1r[j] - p.kappa * lap;
            }
            store_row4(wbase, row, lane, sv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              bm[j] = b0[j];
              b0[j] = bp[j];
            }
          });
        }
      }
      __syncthreads();
      // ---- lam0 = lam1 + s, back in pass-A registers (bit-reversed order) and parked ----
      gather_nat<true>(F, x);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float2 v[8];
        park.load(ch, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = add2(v[i], x[ch * 8 + i]);
          x[ch * 8 + i] = v[i];
        }
        park.store(ch, v);
      }
      tmem_wait_st();
      __syncthreads();  // the natural-layout reads are done before pass A of the next step writes the exchange layout
    }

    // ---- epilogue: lam0 (bit-reversed register order) -> natural layout -> global; coefficient cotangents ----
    {
      float2 y[32];
#pragma unroll
      for (int n = 0; n < 32; ++n) y[n] = x[brev<5>(n)];
      scatter_nat(F, y);
    }
    __syncthreads();
    {
      float* le = p.lam0 + (size_t)env * kRows * kCols;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v[4];
        load_row4(wbase, r0 + i, lane, v);
        *reinterpret_cast<float4*>(le + (r0 + i) * kCols + 4 * lane) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    if (tid < 32) {
      double t = 0.0;
#pragma unroll
      for (int wv = 0; wv < kThreadsR / 32; ++wv) t += S.gacc[wv][tid];
      if (t != 0.0) {
        if (tid < 16) p.gmu[(size_t)env * 16 + tid] += t;
        else p.gmob[(size_t)env * 16 + tid - 16] += t;
      }
    }
    __syncthreads();
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(S.tmem_base));
  }
}

}  // namespace rf
}  // namespace pdeopt
