// Fused K-step semi-implicit Fourier-spectral stepper for 128x128 real fields (sm_100a).
//
// Replaces, for one launch, K iterations of the diffeqsolve loop body of the reference:
//   SemiImplicitFourierSpectral.step        pde_opt/numerics/solvers.py:56-70
//   CahnHilliard2DPeriodic.rhs_fd           pde_opt/numerics/equations/cahn_hilliard.py:89-109
//   AllenCahn2DPeriodic.rhs_fd              pde_opt/numerics/equations/allen_cahn.py:81-84
//   stencils                                pde_opt/numerics/utils/derivatives.py:8-66
// and the observation/reward callbacks of PDEEnv.step (pde_opt/pde_env.py:305-309).
//
// Design (DESIGN.md has the full derivation):
//  * one CTA (512 threads) owns TWO environments packed as one complex field z = u_a + i u_b.
//    The IMEX multiplier 1/(1 + A dt sigma(k)) is real and even, i.e. a real convolution
//    kernel, so filtering z filters u_a and u_b independently: one complex 128x128 FFT pair
//    serves two environments with no Hermitian untangling.
//  * the field never leaves the SM for K steps: 128 KB of shared memory holds it in the
//    "natural" (spatial) layout for the finite-difference RHS and doubles as the exchange
//    buffer of the FFT; the state y that the update y1 = y0 + dt*g needs is parked in
//    tensor memory (TMEM, 64 columns per thread) because registers hold the FFT data.
//  * the 2-D FFT is three register passes (radix 32 | 4x8 | 16x2, all twiddles compile-time
//    except 10 per-thread inter-pass twiddles) with two shared-memory exchanges each way;
//    all exchange patterns are bank-conflict free (tools/fft_decomp_model.py proves it).
//  * the RHS marches down 8 rows per warp with a rolling register window; column
//    neighbours come from warp shuffles (a warp spans a whole periodic row).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pointwise.cuh"
#include "regfft.cuh"
#include "fft128.cuh"

namespace pdeopt {

constexpr int kTabDim = 65;
constexpr int kTabLen = kTabDim * kTabDim;
constexpr int kMaxK = 512;
constexpr int kNCtrl = 8;

enum : int { EQ_CH = 0, EQ_AC = 1 };
enum : int { MODE_FUSED = 0, MODE_RHS_ONLY = 1, MODE_GIVEN_F = 2 };

struct SifsParams {
  const float* y0;
  float* y1;
  int batch;
  int ksteps;
  const float* symbol;  // [kTabLen] folded A*sigma(|kx|,|ky|) (solvers.py:62), or null in rhs-only mode
  const float* ctrl;  // [batch][kNCtrl] or null
  uint8_t* obs;       // [batch][128][128] or null
  float obs_lo, obs_scale;
  float* reward;  // [batch][2] or null
  float* park;    // global parking scratch (only when built with PDEOPT_PARK_GLOBAL)
  const float* f0;  // MODE_GIVEN_F: externally evaluated vector field [batch][128][128]
  int32_t* nonfinite;  // [batch] or null: 1 where y1 of the environment holds a NaN / Inf
  int* work_counter;   // sifs128r: next environment index, zeroed before the launch (dynamic distribution)
  float* traj;         // sifs128r: null, or [ceil(ksteps/save_every)][batch][128][128]: the state at the START of every
  int save_every;      //           save_every-th step of this launch (residuals of the adjoint / tangent rollouts)
  int mode;         // MODE_FUSED / MODE_RHS_ONLY / MODE_GIVEN_F
  float inv_hx, inv_hy, inv_hx2, inv_hy2, kappa;
  float lo_x, lo_y, hx, hy;
  // sifs128r: loop-invariant constants of the RHS phase, both halves equal, filled on the host so that the
  // packed instructions take them straight from the constant bank instead of holding them in registers.
  // S = 1/(2 hx^2) (Cahn-Hilliard) or -1 (Allen-Cahn), see rhs_phase_r.
  float2 rc_ax, rc_ay, rc_a0, rc_ln2p, rc_ln2m, rc_ratio, rc_S;
  PointwiseParams pw;
  float dt[kMaxK];
};

inline void sifs_fill_rhs_consts(SifsParams& p, bool allen_cahn) {
  const float S = allen_cahn ? -1.0f : 0.5f * p.inv_hx * p.inv_hx;
  auto two = [](float v) { return make_float2(v, v); };
  p.rc_S = two(S);
  p.rc_ax = two(-S * p.kappa * p.inv_hx2);
  p.rc_ay = two(-S * p.kappa * p.inv_hy2);
  p.rc_a0 = two(2.0f * S * p.kappa * (p.inv_hx2 + p.inv_hy2));
  p.rc_ln2p = two(0.69314718055994531f * S);
  p.rc_ln2m = two(-0.69314718055994531f * S);
  p.rc_ratio = two((p.inv_hy * p.inv_hy) / (p.inv_hx * p.inv_hx));
}

// ---- TMEM parking of the state (64 x 32-bit columns per thread) --------------------------
#ifndef PDEOPT_PARK_GLOBAL
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float2 (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y),
      "f"(v[4].x), "f"(v[4].y), "f"(v[5].x), "f"(v[5].y), "f"(v[6].x), "f"(v[6].y), "f"(v[7].x), "f"(v[7].y)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float2 (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
        "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y)
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#endif

struct Park {
#ifndef PDEOPT_PARK_GLOBAL
  uint32_t taddr;
#else
  float2* g;  // [32][512] per CTA, element-major so a warp's accesses coalesce
#endif
  __device__ __forceinline__ void store(int chunk, const float2 (&v)[8]) const {
#ifndef PDEOPT_PARK_GLOBAL
    tmem_st16(taddr + chunk * 16, v);
#else
#pragma unroll
    for (int i = 0; i < 8; ++i) g[(chunk * 8 + i) * kThreads + threadIdx.x] = v[i];
#endif
  }
  __device__ __forceinline__ void load(int chunk, float2 (&v)[8]) const {
#ifndef PDEOPT_PARK_GLOBAL
    tmem_ld16(taddr + chunk * 16, v);
#else
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = g[(chunk * 8 + i) * kThreads + threadIdx.x];
#endif
  }
  __device__ __forceinline__ void fence_store() const {
#ifndef PDEOPT_PARK_GLOBAL
    tmem_wait_st();
#endif
  }
};

// ---- small helpers ------------------------------------------------------------------------
__device__ __forceinline__ float2 shfl2(float2 v, int src) {
  return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 f2scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// One row of the natural layout for the S (stencil) mapping: lane l owns columns 4l..4l+3.
__device__ __forceinline__ void load_row(const float2* __restrict__ W, int r, int lane, float2 (&v)[4]) {
  const int p0 = r * kN + ((2 * lane) ^ ((r & 3) << 1));
  const float4 a = *reinterpret_cast<const float4*>(W + p0);
  const float4 b = *reinterpret_cast<const float4*>(W + (p0 ^ 8) + 64);
  v[0] = make_float2(a.x, a.y);
  v[1] = make_float2(a.z, a.w);
  v[2] = make_float2(b.x, b.y);
  v[3] = make_float2(b.z, b.w);
}
__device__ __forceinline__ void store_row(float2* __restrict__ W, int r, int lane, const float2 (&v)[4]) {
  const int p0 = r * kN + ((2 * lane) ^ ((r & 3) << 1));
  *reinterpret_cast<float4*>(W + p0) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
  *reinterpret_cast<float4*>(W + (p0 ^ 8) + 64) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
}

struct EnvCtrl {
  float2 w_off;  // per-env offset of the mu scalar (a, b)
  bool has_bump;
};

// ---- RHS phase: W (natural layout) holds u on entry and f0 = rhs(u) on return ------------
// All arithmetic is packed f32x2 over the (env a, env b) pair.  Relative to the reference's
// expression order (cahn_hilliard.py:89-109) the constant factors 1/2, 1/hx, 1/hy are
// collected into cx = 1/(2 hx^2), cy = 1/(2 hy^2): f = cx*(Gx - Gx_prev) + cy*(Gy - Gy_left)
// with G = (D + D_next) * (mu_next - mu); identical in exact arithmetic, a few ulp in float32.
template <int EQ, int MU, int MOB>
__device__ __forceinline__ void rhs_phase(float2* __restrict__ W, const SifsParams& p, const EnvCtrl& ec,
                                          const float2* __restrict__ gx, const float2* __restrict__ gy) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = warp * 8;
  const int lm1 = (lane + 31) & 31, lp1 = (lane + 1) & 31;

  // Halo rows are read before anybody overwrites them with f0.
  float2 hm2[4], hm1[4], hp8[4], hp9[4];
  load_row(W, (r0 + kN - 1) & (kN - 1), lane, hm1);
  load_row(W, (r0 + 8) & (kN - 1), lane, hp8);
  if (EQ == EQ_CH) {
    load_row(W, (r0 + kN - 2) & (kN - 1), lane, hm2);
    load_row(W, (r0 + 9) & (kN - 1), lane, hp9);
  }
  __syncthreads();

  float2 gyv[4];
  if (ec.has_bump) {
#pragma unroll
    for (int j = 0; j < 4; ++j) gyv[j] = gy[4 * lane + j];
  }
  const float2 ihx2 = splat2(p.inv_hx2), ihy2 = splat2(p.inv_hy2), mkappa = splat2(-p.kappa), m2 = splat2(-2.0f);
  const bool square = p.inv_hx2 == p.inv_hy2;  // uniform
  const float2 mkih2 = splat2(-p.kappa * p.inv_hx2);

  // mu and mobility of one row from its three-row neighbourhood.
  auto mu_row = [&](int rho, const float2 (&um)[4], const float2 (&u0)[4], const float2 (&up)[4], float2 (&mu)[4],
                    float2 (&D)[4]) {
    const float2 uL = shfl2(u0[3], lm1), uR = shfl2(u0[0], lp1);
    float2 gxr = make_float2(0.f, 0.f);
    if (ec.has_bump) gxr = gx[rho & (kN - 1)];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 left = (j == 0) ? uL : u0[j - 1];
      const float2 right = (j == 3) ? uR : u0[j + 1];
      float2 mh;
      mu_mob_pair<MU, MOB>(u0[j], p.pw, ec.w_off, mh, D[j]);
      if (ec.has_bump) mh = fma2(gxr, gyv[j], mh);
      if (square) {
        // hx == hy: -kappa lap(u) = (-kappa/h^2) ((u[i+1] + u[i-1]) + (u[j+1] + u[j-1]) - 4u): 5 packed
        // operations instead of 7 (same value as derivatives.py:8-12 up to the summation order)
        const float2 s4 = add2(add2(up[j], um[j]), add2(right, left));
        mu[j] = fma2(fma2(u0[j], splat2(-4.0f), s4), mkih2, mh);
      } else {
        // derivatives.py:8-12: (u[i+1] - 2u + u[i-1])/hx^2 + (u[j+1] - 2u + u[j-1])/hy^2
        const float2 dxx = add2(fma2(u0[j], m2, up[j]), um[j]);
        const float2 dyy = add2(fma2(u0[j], m2, right), left);
        const float2 lap = fma2(dyy, ihy2, mul2(dxx, ihx2));
        mu[j] = fma2(lap, mkappa, mh);
      }
    }
  };

  if (EQ == EQ_AC) {
    // f = -R(u) * mu   (allen_cahn.py:81-84)
    float2 um[4], u0[4], up[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) um[j] = hm1[j];
    load_row(W, r0, lane, u0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < 7) {
        load_row(W, r0 + i + 1, lane, up);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) up[j] = hp8[j];
      }
      float2 mu[4], R[4], f[4];
      mu_row(r0 + i, um, u0, up, mu, R);
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = mul2(mul2(R[j], splat2(-1.0f)), mu[j]);
      store_row(W, r0 + i, lane, f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        um[j] = u0[j];
        u0[j] = up[j];
      }
    }
    return;
  }

  // Cahn-Hilliard: f = div( D_face * grad_face(mu) )   (cahn_hilliard.py:89-109)
  const float2 cx = splat2(0.5f * p.inv_hx * p.inv_hx), cy = splat2(0.5f * p.inv_hy * p.inv_hy);
  float2 um[4], u0[4], up[4];
  float2 mu_p[4], D_p[4], gx_old[4], dy_p[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    um[j] = hm2[j];
    u0[j] = hm1[j];
  }
  load_row(W, r0, lane, up);
  // it = -1 .. 8  <->  rho = r0 + it : compute mu/D of row rho, emit f of row rho-1.
#pragma unroll
  for (int it = -1; it <= 8; ++it) {
    float2 mu[4], D[4];
    mu_row(r0 + it + kN, um, u0, up, mu, D);
    float2 dy[4];
    if (it >= 0 && it <= 7) {
      const float2 muR = shfl2(mu[0], lp1), DR = shfl2(D[0], lp1);
      float2 g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 mr = (j == 3) ? muR : mu[j + 1];
        const float2 dr = (j == 3) ? DR : D[j + 1];
        g[j] = mul2(add2(D[j], dr), sub2(mr, mu[j]));
      }
      const float2 gL = shfl2(g[3], lm1);
#pragma unroll
      for (int j = 0; j < 4; ++j) dy[j] = sub2(g[j], (j == 0) ? gL : g[j - 1]);
    }
    if (it >= 0) {
      float2 gxn[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) gxn[j] = mul2(add2(D_p[j], D[j]), sub2(mu[j], mu_p[j]));
      if (it >= 1) {
        float2 f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = fma2(sub2(gxn[j], gx_old[j]), cx, mul2(dy_p[j], cy));
        store_row(W, r0 + it - 1, lane, f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) gx_old[j] = gxn[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mu_p[j] = mu[j];
      D_p[j] = D[j];
      if (it >= 0 && it <= 7) dy_p[j] = dy[j];
      um[j] = u0[j];
      u0[j] = up[j];
    }
    // next "up" row: rho + 2 = r0 + it + 2
    if (it + 2 <= 7) {
      load_row(W, r0 + it + 2, lane, up);
    } else if (it + 2 == 8) {
#pragma unroll
      for (int j = 0; j < 4; ++j) up[j] = hp8[j];
    } else if (it + 2 == 9) {
#pragma unroll
      for (int j = 0; j < 4; ++j) up[j] = hp9[j];
    }
  }
}

// ---- spectral filter: f0 (natural layout in W) -> g = Re ifft(fft(f0) / (1 + dt A sigma)) --------
// `mt` is the per-step multiplier table m[kr][fc] = 1 / (N^2 (1 + dt A sigma(|kr|, fc))) with the
// ROW index unfolded (kr = 0..127), so that the 16 rows a thread needs (kr = k1r + 8 brev4(pp)) are
// compile-time offsets from two per-thread base addresses: 32 LDS with immediate offsets, no index
// arithmetic, no division in the step loop (the table is rebuilt only when dt changes).
constexpr int kMRows = kN;
constexpr int kMLen = kMRows * kTabDim;

template <int OFF>
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
  return v;
}

__device__ __forceinline__ void build_multiplier(float* __restrict__ mt, const float* __restrict__ tab, float dt) {
  // solvers.py:62-63 with the inverse-FFT scale folded in
  for (int i = threadIdx.x; i < kMLen; i += kThreads) {
    const int kr = i / kTabDim, fc = i - kr * kTabDim;
    const int fr = kr <= 64 ? kr : 128 - kr;
    mt[i] = __fdividef(1.0f / float(kN * kN), fmaf(dt, tab[fr * kTabDim + fc], 1.0f));
  }
}

__device__ __forceinline__ void spectral_filter(const Fft128& F, uint32_t mt_saddr, float2 (&x)[32]) {
  p1_gather_nat(F.nb, x);
  __syncthreads();  // everybody has read f0 before the buffer is reused as exchange space
  F.forward(x);
  static_for<0, 2>([&](auto bc) {
    constexpr int b = decltype(bc)::value;
    const int kc = F.p3_kc(b);
    const int fc = kc <= 64 ? kc : 128 - kc;
    const uint32_t base = mt_saddr + (uint32_t)(F.m3.k1r * kTabDim + fc) * 4u;
    static_for<0, 16>([&](auto pc) {
      constexpr int pp = decltype(pc)::value;
      const float mval = lds32<brev<4>(pp) * 8 * kTabDim * 4>(base);
      x[b * 16 + pp] = mul2(x[b * 16 + pp], splat2(mval));
    });
  });
  F.inverse(x);
}

// ---- block reductions for the reward epilogue -----------------------------------------------
__device__ __forceinline__ float2 block_sum2(float2 v, float2* red) {
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
  for (int w = 0; w < kThreads / 32; ++w) {
    s.x += red[w].x;
    s.y += red[w].y;
  }
  return s;
}

struct __align__(1024) SifsSmem {
  float2 W[kN * kN];
  float tab[kTabLen + 3];
  float mt[kMLen];
  float2 tw[128];
  float2 gx[kN], gy[kN];
  float2 red[kThreads / 32];
  uint32_t tmem_base;
};

template <int EQ, int MU, int MOB>
__global__ void __launch_bounds__(kThreads, 1) sifs128_kernel(const __grid_constant__ SifsParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SifsSmem& S = *reinterpret_cast<SifsSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int env_a = 2 * blockIdx.x;
  const int env_b = (env_a + 1 < p.batch) ? env_a + 1 : env_a;  // odd batch: duplicate the last env
  const bool b_valid = env_a + 1 < p.batch;

  // ---- setup: TMEM, tables, twiddles, control ----
  Park park;
#ifndef PDEOPT_PARK_GLOBAL
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&S.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
#endif
  if (p.symbol != nullptr)
    for (int i = tid; i < kTabLen; i += kThreads) S.tab[i] = p.symbol[i];
  if (tid < 128) {
    float s, c;
    sincospif(-2.0f * float(tid) / 128.0f, &s, &c);
    S.tw[tid] = make_float2(c, s);
  }
  EnvCtrl ec;
  ec.w_off = make_float2(0.f, 0.f);
  ec.has_bump = false;
  if (p.ctrl != nullptr) {
    const float* ca = p.ctrl + (size_t)env_a * kNCtrl;
    const float* cb = p.ctrl + (size_t)env_b * kNCtrl;
    ec.w_off = make_float2(ca[0], cb[0]);
    ec.has_bump = true;
    if (tid < 2 * kN) {
      // separable Gaussian bump amp*exp(-(x-x0)^2/(2s^2)) * exp(-(y-y0)^2/(2s^2)); amp folded into gx
      const int i = tid & (kN - 1);
      const bool isx = tid < kN;
      float2 out;
      {
        const float pos = isx ? (p.lo_x + (i + 0.5f) * p.hx) : (p.lo_y + (i + 0.5f) * p.hy);
        const float da = pos - (isx ? ca[2] : ca[3]), db = pos - (isx ? cb[2] : cb[3]);
        const float ia = 0.5f / (ca[4] * ca[4]), ib = 0.5f / (cb[4] * cb[4]);
        out.x = (ca[1] != 0.f ? expf(-da * da * ia) : 0.f) * (isx ? ca[1] : 1.0f);
        out.y = (cb[1] != 0.f ? expf(-db * db * ib) : 0.f) * (isx ? cb[1] : 1.0f);
      }
      if (isx) S.gx[i] = out; else S.gy[i] = out;
    }
  }
#ifndef PDEOPT_PARK_GLOBAL
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
  __syncthreads();
#ifndef PDEOPT_PARK_GLOBAL
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  park.taddr = S.tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;
#else
  park.g = reinterpret_cast<float2*>(p.park) + (size_t)blockIdx.x * 32 * kThreads;
#endif

  // ---- prologue: y0 of both envs -> natural layout (S mapping: warp = 8 rows, lane = 4 cols) ----
  {
    const float* ya = p.y0 + (size_t)env_a * kN * kN;
    const float* yb = p.y0 + (size_t)env_b * kN * kN;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      const float4 a = *reinterpret_cast<const float4*>(ya + r * kN + 4 * lane);
      const float4 b = *reinterpret_cast<const float4*>(yb + r * kN + 4 * lane);
      float2 v[4] = {make_float2(a.x, b.x), make_float2(a.y, b.y), make_float2(a.z, b.z), make_float2(a.w, b.w)};
      store_row(S.W, r, lane, v);
    }
  }
  __syncthreads();
  const Fft128 F((uint32_t)__cvta_generic_to_shared(S.W), S.tw);
  float2 x[32];
  p1_gather_nat(F.nb, x);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = x[ch * 8 + i];
    park.store(ch, v);
  }
  park.fence_store();
  __syncthreads();

  // ---- K fused steps ----
  if (p.mode == MODE_RHS_ONLY) {
    // eq.rhs(state, t): emit f0 = rhs(y0) instead of stepping (cahn_hilliard.py:89-109).
    rhs_phase<EQ, MU, MOB>(S.W, p, ec, S.gx, S.gy);
    __syncthreads();
  }
  const uint32_t mt_saddr = (uint32_t)__cvta_generic_to_shared(S.mt);
  float dt_tab = __int_as_float(0x7fc00000);  // NaN: no table yet
  for (int k = 0; k < ((p.mode == MODE_RHS_ONLY) ? 0 : p.ksteps); ++k) {
    if (p.dt[k] != dt_tab) {
      // rebuilt only when the step length changes; the barriers of the RHS / load phase below order
      // these writes before the reads in spectral_filter, and the previous step's reads before them
      dt_tab = p.dt[k];
      build_multiplier(S.mt, S.tab, dt_tab);
    }
    if (p.mode == MODE_GIVEN_F) {
      // unfused vector field (terms.vf evaluated by the caller, solvers.py:59): load f0 instead
      const float* fa = p.f0 + (size_t)env_a * kN * kN;
      const float* fb = p.f0 + (size_t)env_b * kN * kN;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 8 + i;
        const float4 a = *reinterpret_cast<const float4*>(fa + r * kN + 4 * lane);
        const float4 b = *reinterpret_cast<const float4*>(fb + r * kN + 4 * lane);
        float2 v[4] = {make_float2(a.x, b.x), make_float2(a.y, b.y), make_float2(a.z, b.z), make_float2(a.w, b.w)};
        store_row(S.W, r, lane, v);
      }
    } else {
      rhs_phase<EQ, MU, MOB>(S.W, p, ec, S.gx, S.gy);
    }
    __syncthreads();
    const float dt = p.dt[k];
    spectral_filter(F, mt_saddr, x);
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
    park.fence_store();
    __syncthreads();  // all exchange-layout reads are done before the natural layout is rewritten
    p1_scatter_nat(F.nb, x);
    __syncthreads();
  }

  // ---- epilogue: y1 to global (coalesced), optional uint8 observation and (mean, var) ----
  {
    float* ya = p.y1 + (size_t)env_a * kN * kN;
    float* yb = p.y1 + (size_t)env_b * kN * kN;
    float2 sum = make_float2(0.f, 0.f);
    float2 rows[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      load_row(S.W, r, lane, rows[i]);
      const float2* v = rows[i];
      *reinterpret_cast<float4*>(ya + r * kN + 4 * lane) = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
      if (b_valid) *reinterpret_cast<float4*>(yb + r * kN + 4 * lane) = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sum.x += v[j].x;
        sum.y += v[j].y;
      }
      if (p.obs != nullptr) {
        uint32_t pa = 0, pb = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float qa = rintf(__saturatef((v[j].x - p.obs_lo) * p.obs_scale) * 255.0f);
          const float qb = rintf(__saturatef((v[j].y - p.obs_lo) * p.obs_scale) * 255.0f);
          pa |= (uint32_t)qa << (8 * j);
          pb |= (uint32_t)qb << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(p.obs + (size_t)env_a * kN * kN + r * kN + 4 * lane) = pa;
        if (b_valid) *reinterpret_cast<uint32_t*>(p.obs + (size_t)env_b * kN * kN + r * kN + 4 * lane) = pb;
      }
    }
    if (p.reward != nullptr) {
      const float inv_n = 1.0f / float(kN * kN);
      const float2 tot = block_sum2(sum, S.red);
      const float2 mean = make_float2(tot.x * inv_n, tot.y * inv_n);
      float2 sq = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float da = rows[i][j].x - mean.x, db = rows[i][j].y - mean.y;
          sq.x = fmaf(da, da, sq.x);
          sq.y = fmaf(db, db, sq.y);
        }
      const float2 tsq = block_sum2(sq, S.red);
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
#ifndef PDEOPT_PARK_GLOBAL
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(S.tmem_base));
  }
#endif
}

}  // namespace pdeopt
