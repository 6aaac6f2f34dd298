// Fused K-step semi-implicit Fourier-spectral stepper for 128x128 real fields, ONE environment per
// 256-thread CTA, two CTAs per SM (sm_100a).
//
// Replaces, for one launch, K iterations of the diffeqsolve loop body of the reference:
//   SemiImplicitFourierSpectral.step        pde_opt/numerics/solvers.py:56-70
//   CahnHilliard2DPeriodic.rhs_fd           pde_opt/numerics/equations/cahn_hilliard.py:89-109
//   AllenCahn2DPeriodic.rhs_fd              pde_opt/numerics/equations/allen_cahn.py:81-84
//   stencils                                pde_opt/numerics/utils/derivatives.py:8-66
// and the observation/reward callbacks of PDEEnv.step (pde_opt/pde_env.py:305-309).
//
// Why one REAL field per CTA (round-1 kernel: two environments packed as one complex field per
// 512-thread CTA, 128 KB of shared memory, one CTA per SM): all warps of a CTA sit in the same
// barrier-delimited phase, so the FP32-pipe phases (butterflies, stencil) and the shared-memory
// phases (exchanges) of that design ADD.  Here the field is 64 KB (z[r][m] = u[r][2m] + i u[r][2m+1]),
// two independent CTAs share an SM and are in different phases, so one's exchanges overlap the
// other's arithmetic; no environment shares any arithmetic with another one (a non-finite state
// stays in its own environment).  rfft128.cuh holds the transform and the closed-form filter.
//
// Per CTA: 64 KB field buffer (natural layout for the stencil, exchange layout for the passes),
// 33 KB filter table per distinct dt, y0 of the step parked in 128 TMEM columns, 32 complex
// registers per thread.  The CTA is persistent over environments (grid = min(batch, 2 x SMs)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rfft128.cuh"
#include "sifs128.cuh"  // SifsParams, EnvCtrl, modes, TMEM helpers, pointwise closures

// Ablation hooks for timing experiments (tools/_exp builds only; results are wrong without barriers).
#ifdef PDEOPT_EXP_NO_BARRIERS
#define PDEOPT_EXP_SYNC() __syncwarp()
#else
#define PDEOPT_EXP_SYNC() __syncthreads()
#endif

namespace pdeopt {
namespace rf {

struct __align__(1024) RSmem {
  float2 W[kRows * kH];
  float4 T[kTRows * kTCols];
  float2 twb[8 * 16];
  float2 tw64[32];
  float2 sc[32];
  float gx[kRows], gy[kCols];
  float2 red[kThreadsR / 32];
  uint32_t tmem_base;
  int next_env;
};

struct ParkR {
  uint32_t taddr;
  __device__ __forceinline__ void store(int chunk, const float2 (&v)[8]) const { tmem_st16(taddr + chunk * 16, v); }
  __device__ __forceinline__ void load(int chunk, float2 (&v)[8]) const { tmem_ld16(taddr + chunk * 16, v); }
};

// S (stencil) map: warp = 16 rows, lane l = columns 4l..4l+3 = complex slots 2l, 2l+1 (one 16-byte access).
__device__ __forceinline__ uint32_t srow_addr(uint32_t wbase, int r, int lane) {
  return wbase + (uint32_t)(r * 512 + ((lane * 16) ^ ((r & 7) << 4)));
}
__device__ __forceinline__ void load_srow(uint32_t wbase, int r, int lane, float2 (&v)[2]) {
  const float4 a = ld4<0>(srow_addr(wbase, r, lane));
  v[0] = make_float2(a.x, a.y);
  v[1] = make_float2(a.z, a.w);
}
__device__ __forceinline__ void store_srow(uint32_t wbase, int r, int lane, const float2 (&v)[2]) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow_addr(wbase, r, lane)), "f"(v[0].x), "f"(v[0].y),
               "f"(v[1].x), "f"(v[1].y)
               : "memory");
}

__device__ __forceinline__ float shf(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ---- RHS phase: W (natural layout) holds u on entry and f0 = rhs(u) on return ------------------
// Arithmetic is packed f32x2 over the column pairs (2m, 2m+1) wherever both operands are aligned
// pairs; differences and sums of column NEIGHBOURS straddle the pairs and are scalar.  Relative to
// the reference's expression order (cahn_hilliard.py:89-109) the constant factors 1/2, 1/hx, 1/hy
// are collected into cx = 1/(2 hx^2), cy = 1/(2 hy^2): f = cx (Gx - Gx_prev) + cy (Gy - Gy_left),
// G = (D + D_next)(mu_next - mu); identical in exact arithmetic, a few ulp in float32.
template <int EQ, int MU, int MOB>
__device__ __forceinline__ void rhs_phase_r(uint32_t wbase, const SifsParams& p, float w_off, bool has_bump,
                                            const float* __restrict__ gx, const float* __restrict__ gy) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = warp * 16;
  const int lm1 = (lane + 31) & 31, lp1 = (lane + 1) & 31;

  // halo rows are read before anybody overwrites them with f0
  float2 hm2[2], hm1[2], hp0[2], hp1[2];
  load_srow(wbase, (r0 + kRows - 1) & (kRows - 1), lane, hm1);
  load_srow(wbase, (r0 + 16) & (kRows - 1), lane, hp0);
  if (EQ == EQ_CH) {
    load_srow(wbase, (r0 + kRows - 2) & (kRows - 1), lane, hm2);
    load_srow(wbase, (r0 + 17) & (kRows - 1), lane, hp1);
  }
  __syncthreads();

  float2 gyv[2];
  if (has_bump) {
    const float4 g4 = *reinterpret_cast<const float4*>(gy + 4 * lane);
    gyv[0] = make_float2(g4.x, g4.y);
    gyv[1] = make_float2(g4.z, g4.w);
  }
  // Everything that feeds mu is pre-multiplied by one constant S so that the flux needs no trailing scale:
  //   Cahn-Hilliard S = cx = 1/(2 hx^2):  f = (Gx - Gx_prev) + (cy/cx) (Gy - Gy_left),  G = (D + D_next)(S mu_next - S mu)
  //   Allen-Cahn    S = -1:               f = R * (S mu)
  // S mu = S ax (u[i+1] + u[i-1]) + S ay (u[j+1] + u[j-1]) + S a0 u + S mu_h with ax = -kappa/hx^2, ay = -kappa/hy^2,
  // a0 = 2 kappa (1/hx^2 + 1/hy^2): one packed add and three packed FMAs per pair, the constants folded on the
  // way in (derivatives.py:8-12 and cahn_hilliard.py:89-109 up to the summation order and a common factor; the
  // same value to a few ulp of the largest term).  The control bump table gx is pre-scaled by S by the caller.
  const float2 S2 = p.rc_S, ax = p.rc_ax, ay = p.rc_ay, a0 = p.rc_a0, ln2p = p.rc_ln2p, ln2m = p.rc_ln2m;
  const float2 woff2 = splat2(w_off);
  // log family: S mu_h + S a0 c = (S a0 - 2 S w) c + S w + S ln2 lg2(c) - S ln2 lg2(1 - c): three chained FMAs
  const float wlog = p.pw.mu_coef[0] + w_off;
  const float2 wl2 = splat2(S2.x * wlog), lin2 = splat2(fmaf(-2.0f * S2.x, wlog, a0.x));

  // S mu and the mobility of one row from its three-row neighbourhood
  auto mu_row = [&](int rho, const float2 (&um)[2], const float2 (&u0)[2], const float2 (&up)[2], float2 (&mu)[2],
                    float2 (&D)[2]) {
    const float uL = shf(u0[1].y, lm1), uR = shf(u0[0].x, lp1);
    float2 gxr = make_float2(0.f, 0.f);
    if (has_bump) gxr = splat2(gx[rho & (kRows - 1)]);
    // left + right column neighbours: the operands straddle the aligned pairs.  Scalar adds cost the same
    // FP32-pipe cycles as two packed adds and save the register moves that assembling the shifted pairs
    // (uL, c0), (c1, c2), (c3, uR) would take (issue slots); PDEOPT_RHS_PACKED_SHIFTS keeps the packed form.
    float2 s[2];
#ifdef PDEOPT_RHS_PACKED_SHIFTS
    const float2 q = make_float2(u0[0].y, u0[1].x);
    s[0] = add2(make_float2(uL, u0[0].x), q);
    s[1] = add2(q, make_float2(u0[1].y, uR));
#else
    s[0] = make_float2(uL + u0[0].y, u0[0].x + u0[1].x);
    s[1] = make_float2(u0[0].y + u0[1].y, u0[1].x + uR);
#endif
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float2 mh;
      if constexpr (MU == MU_LOG && (MOB == MOB_DEGENERATE || MOB == MOB_CONST)) {
        const float2 c = u0[j];
        const float2 sm = sub2(splat2(1.0f), c);
        mh = fma2(c, lin2, wl2);
        mh = fma2(make_float2(lg2_fast(sm.x), lg2_fast(sm.y)), ln2m, mh);
        mh = fma2(make_float2(lg2_fast(c.x), lg2_fast(c.y)), ln2p, mh);
        D[j] = (MOB == MOB_DEGENERATE) ? mul2(sm, c) : splat2(p.pw.mob_coef[0]);
      } else {
        mu_mob_pair<MU, MOB>(u0[j], p.pw, woff2, mh, D[j]);
        mh = fma2(a0, u0[j], mul2(mh, S2));
      }
      if (has_bump) mh = fma2(gxr, gyv[j], mh);
      mu[j] = fma2(ax, add2(up[j], um[j]), fma2(ay, s[j], mh));
    }
  };

  if (EQ == EQ_AC) {
    // f = -R(u) * mu   (allen_cahn.py:81-84)
    float2 um[2], u0[2], up[2];
    um[0] = hm1[0];
    um[1] = hm1[1];
    load_srow(wbase, r0, lane, u0);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < 15) {
        load_srow(wbase, r0 + i + 1, lane, up);
      } else {
        up[0] = hp0[0];
        up[1] = hp0[1];
      }
      float2 mu[2], R[2], f[2];
      mu_row(r0 + i, um, u0, up, mu, R);
#pragma unroll
      for (int j = 0; j < 2; ++j) f[j] = mul2(R[j], mu[j]);
      store_srow(wbase, r0 + i, lane, f);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        um[j] = u0[j];
        u0[j] = up[j];
      }
    }
    return;
  }

  // Cahn-Hilliard: f = div( D_face * grad_face(mu) )   (cahn_hilliard.py:89-109)
  const float2 ratio = p.rc_ratio;  // cy / cx
  float2 um[2], u0[2], up[2];
  float2 mu_p[2], D_p[2], gx_old[2], dy_p[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    um[j] = hm2[j];
    u0[j] = hm1[j];
  }
  load_srow(wbase, r0, lane, up);
  // it = -1 .. 16  <->  rho = r0 + it : compute mu/D of row rho, emit f of row rho-1
#pragma unroll
  for (int it = -1; it <= 16; ++it) {
    float2 mu[2], D[2];
    mu_row(r0 + it + kRows, um, u0, up, mu, D);
    float2 dy[2];
    if (it >= 0 && it <= 15) {
      const float muR = shf(mu[0].x, lp1), DR = shf(D[0].x, lp1);
      float2 g[2];
      const float2 qD = make_float2(D[0].y, D[1].x), qm = make_float2(mu[0].y, mu[1].x);
      g[0] = mul2(add2(D[0], qD), sub2(qm, mu[0]));
      g[1] = mul2(add2(D[1], make_float2(D[1].y, DR)), sub2(make_float2(mu[1].y, muR), mu[1]));
      const float gL = shf(g[1].y, lm1);
      dy[0] = sub2(g[0], make_float2(gL, g[0].x));
#ifdef PDEOPT_RHS_PACKED_SHIFTS
      dy[1] = sub2(g[1], make_float2(g[0].y, g[1].x));
#else
      dy[1] = make_float2(g[1].x - g[0].y, g[1].y - g[1].x);
#endif
    }
    if (it >= 0) {
      float2 gxn[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) gxn[j] = mul2(add2(D_p[j], D[j]), sub2(mu[j], mu_p[j]));
      if (it >= 1) {
        float2 f[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) f[j] = fma2(dy_p[j], ratio, sub2(gxn[j], gx_old[j]));
        store_srow(wbase, r0 + it - 1, lane, f);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) gx_old[j] = gxn[j];
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      mu_p[j] = mu[j];
      D_p[j] = D[j];
      if (it >= 0 && it <= 15) dy_p[j] = dy[j];
      um[j] = u0[j];
      u0[j] = up[j];
    }
    // next "up" row: rho + 2 = r0 + it + 2
    if (it + 2 <= 15) {
      load_srow(wbase, r0 + it + 2, lane, up);
    } else if (it + 2 == 16) {
      up[0] = hp0[0];
      up[1] = hp0[1];
    } else if (it + 2 == 17) {
      up[0] = hp1[0];
      up[1] = hp1[1];
    }
  }
}

__device__ __forceinline__ void build_table_r(float4* __restrict__ T, const float* __restrict__ tab,
                                              const float2* __restrict__ sc, float dt) {
  // solvers.py:62-63 with the inverse-FFT scale and the real-transform untangling folded in
  for (int i = threadIdx.x; i < kTRows * kTCols; i += kThreadsR) T[i] = filter_entry(tab, sc, i >> 5, i & 31, dt);
}

__device__ __forceinline__ float2 block_sum2_r(float2 v, float2* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < kThreadsR / 32; ++w) {
    s.x += red[w].x;
    s.y += red[w].y;
  }
  return s;
}

template <int EQ, int MU, int MOB>
__global__ void __launch_bounds__(kThreadsR, 2) sifs128r_kernel(const __grid_constant__ SifsParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  RSmem& S = *reinterpret_cast<RSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- once per CTA: TMEM, twiddles ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(
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
  ParkR park;
  park.taddr = S.tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;

  const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(S.W);
  const RFft F(wbase, (uint32_t)__cvta_generic_to_shared(S.T), tid);
  float dt_tab = __int_as_float(0x7fc00000);  // NaN: no table yet

#ifdef PDEOPT_SKEW_CYCLES
  if (blockIdx.x >= gridDim.x / 2) {  // experiment: start half of the CTAs out of phase
    const long long t0 = clock64();
    while (clock64() - t0 < PDEOPT_SKEW_CYCLES) {}
  }
#endif
  // Environments are handed out dynamically (one atomic per environment): the two CTAs of an SM do not
  // progress at the same rate (the warp scheduler favours one of them), so a static split would leave
  // the favoured CTA idle at the end of the launch while the other one finishes alone.
  for (int env = blockIdx.x; env < p.batch;) {
    // the NEXT environment is claimed now, so that its state can be pulled into L2 while this one is stepped
    if (tid == 0) S.next_env = (int)gridDim.x + atomicAdd(p.work_counter, 1);
    // ---- per-environment control (pde_env.py:274-286; our definition) ----
    float w_off = 0.f;
    bool has_bump = false;
    if (p.ctrl != nullptr) {
      const float* ce = p.ctrl + (size_t)env * kNCtrl;
      w_off = ce[0];
      has_bump = true;
      // separable Gaussian bump amp*exp(-(x-x0)^2/(2s^2)) * exp(-(y-y0)^2/(2s^2)); amp folded into gx
      const int i = tid & (kRows - 1);
      const bool isx = tid < kRows;
      const float pos = isx ? (p.lo_x + (i + 0.5f) * p.hx) : (p.lo_y + (i + 0.5f) * p.hy);
      const float dd = pos - (isx ? ce[2] : ce[3]);
      const float iw = 0.5f / (ce[4] * ce[4]);
      // gx carries the amplitude and the constant S of rhs_phase_r (Cahn-Hilliard: 1/(2 hx^2); Allen-Cahn: -1)
      const float sx = ce[1] * p.rc_S.x;
      const float v = (ce[1] != 0.f ? expf(-dd * dd * iw) : 0.f) * (isx ? sx : 1.0f);
      if (isx) S.gx[i] = v; else S.gy[i] = v;
    }
    // ---- prologue: y0 -> natural layout (S map) ----
    {
      const float* ye = p.y0 + (size_t)env * kRows * kCols;
      float4 a[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __ldg(reinterpret_cast<const float4*>(ye + (warp * 16 + i) * kCols + 4 * lane));
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float2 v[2] = {make_float2(a[i].x, a[i].y), make_float2(a[i].z, a[i].w)};
        store_srow(wbase, warp * 16 + i, lane, v);
      }
    }
    __syncthreads();
    const int next_env = S.next_env;
#ifndef PDEOPT_NO_L2_PREFETCH
    if (next_env < p.batch) {
      // 64 KB = 512 lines of 128 B, two per thread: the prologue loads of the next environment hit L2, which
      // matters when few steps are fused per launch (K = 1: load, step and store of a CTA do not overlap)
      const char* nx = reinterpret_cast<const char*>(p.y0 + (size_t)next_env * kRows * kCols) + tid * 256;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128));
    }
#endif
    float2 x[32];
    gather_nat(F, x);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = x[ch * 8 + i];
      park.store(ch, v);
    }
    tmem_wait_st();
    __syncthreads();

    if (p.mode == MODE_RHS_ONLY) {
      // eq.rhs(state, t): emit f0 = rhs(y0) instead of stepping (cahn_hilliard.py:89-109)
      rhs_phase_r<EQ, MU, MOB>(wbase, p, w_off, has_bump, S.gx, S.gy);
      __syncthreads();
    }
    for (int k = 0; k < ((p.mode == MODE_RHS_ONLY) ? 0 : p.ksteps); ++k) {
      const float dt = p.dt[k];
      if (dt != dt_tab) {
        // rebuilt only when the step length changes; the barriers of the phases below order these
        // writes before the reads of pass C, and the previous step's reads before them
        dt_tab = dt;
        build_table_r(S.T, p.symbol, S.sc, dt);
      }
      if (p.traj != nullptr && (k % p.save_every) == 0) {
        // checkpoint for the adjoint / tangent rollouts: the natural layout holds the state at the start of the
        // step; every warp copies the rows it is about to overwrite (no extra barrier)
        float* te = p.traj + ((size_t)(k / p.save_every) * p.batch + env) * kRows * kCols;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 v[2];
          load_srow(wbase, warp * 16 + i, lane, v);
          *reinterpret_cast<float4*>(te + (warp * 16 + i) * kCols + 4 * lane) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        }
      }
      if (p.mode == MODE_GIVEN_F) {
        // unfused vector field (terms.vf evaluated by the caller, solvers.py:59): load f0 instead
        const float* fe = p.f0 + (size_t)env * kRows * kCols;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(fe + (warp * 16 + i) * kCols + 4 * lane));
          float2 v[2] = {make_float2(a.x, a.y), make_float2(a.z, a.w)};
          store_srow(wbase, warp * 16 + i, lane, v);
        }
      } else {
#ifndef PDEOPT_EXP_NO_RHS
        rhs_phase_r<EQ, MU, MOB>(wbase, p, w_off, has_bump, S.gx, S.gy);
#endif
      }
      PDEOPT_EXP_SYNC();
      gather_nat<true>(F, x);
      PDEOPT_EXP_SYNC();  // everybody has read f0 before the buffer is reused as exchange space
      passA_fwd(F, x);
      PDEOPT_EXP_SYNC();
      passB_fwd(F, S.twb, S.tw64, x);
      PDEOPT_EXP_SYNC();
      passC_filter(F, x);
      PDEOPT_EXP_SYNC();
      passB_inv(F, S.twb, S.tw64, x);
      PDEOPT_EXP_SYNC();
      passA_inv(F, x);
      // y1 = y0 + dt * g   (solvers.py:63); y0 comes back from the parking space
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float2 v[8];
        park.load(ch, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = fma2(x[ch * 8 + i], splat2(dt), v[i]);
          x[ch * 8 + i] = v[i];
        }
        park.store(ch, v);
      }
      tmem_wait_st();
      PDEOPT_EXP_SYNC();  // all exchange-layout reads are done before the natural layout is rewritten
      scatter_nat(F, x);
      PDEOPT_EXP_SYNC();
    }

    // ---- epilogue: y1 to global (coalesced), optional uint8 observation, (mean, var), non-finite flag ----
    {
      float* ye = p.y1 + (size_t)env * kRows * kCols;
      float sum = 0.f;
      float2 rows[16][2];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = warp * 16 + i;
        load_srow(wbase, r, lane, rows[i]);
        const float2* v = rows[i];
        *reinterpret_cast<float4*>(ye + r * kCols + 4 * lane) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        sum += (v[0].x + v[0].y) + (v[1].x + v[1].y);
        if (p.obs != nullptr) {
          const float q[4] = {v[0].x, v[0].y, v[1].x, v[1].y};
          uint32_t pa = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            pa |= (uint32_t)rintf(__saturatef((q[j] - p.obs_lo) * p.obs_scale) * 255.0f) << (8 * j);
          *reinterpret_cast<uint32_t*>(p.obs + (size_t)env * kRows * kCols + r * kCols + 4 * lane) = pa;
        }
      }
      if (p.reward != nullptr || p.nonfinite != nullptr) {
        const float inv_n = 1.0f / float(kRows * kCols);
        float sq0 = 0.f;
        const float2 tot = block_sum2_r(make_float2(sum, 0.f), S.red);
        const float mean = tot.x * inv_n;
        if (p.reward != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float da = rows[i][j].x - mean, db = rows[i][j].y - mean;
              sq0 = fmaf(da, da, sq0);
              sq0 = fmaf(db, db, sq0);
            }
          const float2 tsq = block_sum2_r(make_float2(sq0, 0.f), S.red);
          if (tid == 0) {
            p.reward[2 * env] = mean;
            p.reward[2 * env + 1] = tsq.x * inv_n;
          }
        }
        // a NaN or Inf anywhere in the field makes the sum non-finite (pde_model.py:131 `throw` semantics)
        if (p.nonfinite != nullptr && tid == 0) p.nonfinite[env] = (fabsf(tot.x) <= 3.0e38f) ? 0 : 1;
      }
    }
    __syncthreads();  // the field buffer and the control tables are reused by the next environment
    env = next_env;
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(S.tmem_base));
  }
}

}  // namespace rf
}  // namespace pdeopt
