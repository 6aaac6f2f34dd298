// Fused K-step advection-diffusion stepper for 128x128 real fields and its hand-written
// discrete adjoint (sm_100a).
//
// The equation is the one recovered in SURVEY F6 (the reference imports AdvectionDiffusion2D in
// notebooks/run_advection_diffusion.ipynb cell 0 but the file is absent from the tree):
//     du/dt = -d/dx(vx u) - d/dy(vy u) + D lap(u),    Fourier-spectral derivatives,
//     v = p0 * grad exp(-r^2 / (2 p1)) about a controlled centre (notebook cell 2),
// stepped by SemiImplicitFourierSpectral.step (pde_opt/numerics/solvers.py:56-70) with the
// symbol sigma = D (2 pi)^2 |k|^2:
//     y1 = y0 + dt * Re ifft( fft(f(y0)) / (1 + A dt sigma) ).
// Because every operator after the pointwise products v*u is a Fourier multiplier, the step is
//     y1 = u + dt * F^-1[ m ( -i kx F[vx u] - i ky F[vy u] - L F[u] ) ],  m = 1/(1 + A dt sigma),
// with L = D (2 pi)^2 |k|^2 and the odd multipliers zeroed on the Nyquist lines (that is what
// the reference's `.real` of the full complex transform amounts to for real fields).  All
// multipliers are spectra of real kernels, so two environments ride in one complex field
// z = u_a + i u_b exactly as in sifs128.cuh.
//
// Backward (the custom_vjp of the rollout, replacing diffrax's RecursiveCheckpointAdjoint driven
// from pde_model.py:226-323): with w = G lam1 (G = F^-1 m F, self-adjoint),
//     lam0 = F^-1[(1 - dt L m) F lam1] + dt vx F^-1[i kx m F lam1] + dt vy F^-1[i ky m F lam1]
//     dLoss/dtheta += dt <d/dx w, u dvx/dtheta> + dt <d/dy w, u dvy/dtheta>,  theta in (cx, cy, p0, p1)
// where u is the state at the start of the step, saved by the forward kernel (the whole
// trajectory fits in HBM: 500 steps x 512 envs x 64 KB = 16 GiB of the 180 GB).
//
// On-chip state: registers hold the FFT working set, TMEM holds two parked fields (512 columns:
// the whole tensor memory of the SM, used as a 256 KB register file extension).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft128.cuh"
#include "sifs128.cuh"
#include "spec_util.cuh"

namespace pdeopt {

constexpr int kAdCtrl = 4;  // (cx, cy, p0, p1) per control segment

struct AdParams {
  const float* y0;   // fwd: [batch][128][128] initial state; bwd: cotangent of the final state
  float* y1;         // fwd: final state; bwd: cotangent of the initial state
  float* traj;       // fwd: if non-null, state at the START of step k -> traj[k*traj_stride ...], [pair][32][512] float2
  const float* traj_in;  // bwd: the same array
  long long traj_stride;
  int batch, ksteps;
  const float* tabA;  // [65*65] A * sigma                (denominator of the IMEX filter)
  const float* tabL;  // [65*65] D (2 pi)^2 |k|^2         (explicit diffusion term of the RHS)
  const float* kx;    // [128] 2 pi kx (imag part of two_pi_i_kx), 0 at the Nyquist index
  const float* ky;    // [128] 2 pi ky, 0 at the Nyquist index
  const float* ctrl;  // [batch][nseg][4] (cx, cy, p0, p1), piecewise constant over `hold` steps
  float* gctrl;       // bwd: [batch][nseg][4], accumulated (+=)
  int nseg, hold, step0;  // control segment of local step k: min((step0 + k) / hold, nseg - 1)
  float lo_x, lo_y, hx, hy;
  float dt[kMaxK];
};

struct __align__(1024) AdSmem {
  float2 W[kN * kN];
  float tabA[kTabLen + 3];
  float tabL[kTabLen + 3];
  float kx[kN], ky[kN];
  float2 tw[128];
  float2 ax[kN], ex[kN], dxv[kN];            // row (x) tables: ax = -(p0/p1) dx ex
  float2 ay[kN], ey[kN];                     // column (y) tables: ay = -(p0/p1) dy ey
  float2 dyey[kN], dy2ey[kN], dy3ey[kN];     // bwd: dy^j ey
  float2 red[kThreads / 32][4];
  float2 segc[4];             // bwd: (p0, 1/p1, p0/p1) of the current segment for the (a, b) pair
  float2 gacc[4][kThreads];   // bwd: per-thread cotangent accumulators (cx, cy, p0, p1), kept out of the register file
  uint32_t tmem_base;
};


// Separable velocity tables of one control segment for the (a, b) environment pair.
__device__ __forceinline__ void ad_tables(AdSmem& S, const AdParams& p, int env_a, int env_b, int seg) {
  const int tid = threadIdx.x;
  if (tid < 2 * kN) {
    const float* ca = p.ctrl + ((size_t)env_a * p.nseg + seg) * kAdCtrl;
    const float* cb = p.ctrl + ((size_t)env_b * p.nseg + seg) * kAdCtrl;
    const int i = tid & (kN - 1);
    const bool isx = tid < kN;
    const float pos = isx ? (p.lo_x + (i + 0.5f) * p.hx) : (p.lo_y + (i + 0.5f) * p.hy);
    const float2 d = make_float2(pos - (isx ? ca[0] : ca[1]), pos - (isx ? cb[0] : cb[1]));
    const float2 p0 = make_float2(ca[2], cb[2]), p1 = make_float2(ca[3], cb[3]);
    const float2 e = make_float2(expf(-d.x * d.x / (2.0f * p1.x)), expf(-d.y * d.y / (2.0f * p1.y)));
    const float2 a = make_float2(p0.x * (-d.x / p1.x * e.x), p0.y * (-d.y / p1.y * e.y));
    if (isx) {
      S.ax[i] = a;
      S.ex[i] = e;
      S.dxv[i] = d;
    } else {
      S.ay[i] = a;
      S.ey[i] = e;
      const float2 de = f2mul(d, e);
      S.dyey[i] = de;
      const float2 d2e = f2mul(d, de);
      S.dy2ey[i] = d2e;
      S.dy3ey[i] = f2mul(d, d2e);
    }
  }
}

struct AdCommon {
  Park park0, park1;
  int env_a, env_b;
  bool b_valid;
};

// TMEM allocation (all 512 columns), constant tables, twiddles.  Ends with a barrier.
__device__ __forceinline__ void ad_setup(AdSmem& S, const AdParams& p, AdCommon& C) {
  const int tid = threadIdx.x, warp = tid >> 5;
  C.env_a = 2 * blockIdx.x;
  C.env_b = (C.env_a + 1 < p.batch) ? C.env_a + 1 : C.env_a;
  C.b_valid = C.env_a + 1 < p.batch;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&S.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < kTabLen; i += kThreads) {
    S.tabA[i] = p.tabA[i];
    S.tabL[i] = p.tabL[i];
  }
  if (tid < kN) {
    S.kx[tid] = p.kx[tid];
    S.ky[tid] = p.ky[tid];
    float s, c;
    sincospif(-2.0f * float(tid) / 128.0f, &s, &c);
    S.tw[tid] = make_float2(c, s);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  C.park0.taddr = S.tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;
  C.park1.taddr = C.park0.taddr + 256u;
}

__device__ __forceinline__ void ad_teardown(AdSmem& S) {
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(S.tmem_base));
}

// pair field [env_a | env_b] in global memory -> natural layout in W
__device__ __forceinline__ void ad_load_pair(float2* W, const float* ya, const float* yb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp * 8 + i;
    const float4 a = *reinterpret_cast<const float4*>(ya + r * kN + 4 * lane);
    const float4 b = *reinterpret_cast<const float4*>(yb + r * kN + 4 * lane);
    float2 v[4] = {make_float2(a.x, b.x), make_float2(a.y, b.y), make_float2(a.z, b.z), make_float2(a.w, b.w)};
    store_row(W, r, lane, v);
  }
}
__device__ __forceinline__ void ad_store_pair(const float2* W, float* ya, float* yb, bool b_valid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp * 8 + i;
    float2 v[4];
    load_row(W, r, lane, v);
    *reinterpret_cast<float4*>(ya + r * kN + 4 * lane) = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
    if (b_valid) *reinterpret_cast<float4*>(yb + r * kN + 4 * lane) = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
  }
}

__device__ __forceinline__ int ad_seg(const AdParams& p, int k) {
  const int s = (p.step0 + k) / p.hold;
  return s < p.nseg - 1 ? s : p.nseg - 1;
}

// =============================================================================================
// forward
// =============================================================================================
__global__ void __launch_bounds__(kThreads, 1) ad128_fwd_kernel(const __grid_constant__ AdParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  AdSmem& S = *reinterpret_cast<AdSmem*>(smem_raw);
  AdCommon C;
  ad_setup(S, p, C);
  const size_t oa = (size_t)C.env_a * kN * kN, ob = (size_t)C.env_b * kN * kN;

  ad_load_pair(S.W, p.y0 + oa, p.y0 + ob);
  __syncthreads();
  const Fft128 F((uint32_t)__cvta_generic_to_shared(S.W), S.tw);
  float2 x[32];
  p1_gather_nat(F.nb, x);
  __syncthreads();
  const int r = F.p1_row();
  const int n2c = F.m1.n2c;

  int cur_seg = -1;
  float2 axr = make_float2(0.f, 0.f), exr = axr;
  for (int k = 0; k < p.ksteps; ++k) {
    const int seg = ad_seg(p, k);
    if (seg != cur_seg) {  // uniform across the CTA
      __syncthreads();
      ad_tables(S, p, C.env_a, C.env_b, seg);
      __syncthreads();
      axr = S.ax[r];
      exr = S.ex[r];
      cur_seg = seg;
    }
    if (p.traj != nullptr) {
      // save the state at the start of the step (what the adjoint needs) straight from the registers:
      // the trajectory is an internal buffer, so it is kept in the P1 register arrangement
      // ([pair][n][thread] float2), written and read back with coalesced 8-byte accesses
      float2* t = reinterpret_cast<float2*>(p.traj + (size_t)k * p.traj_stride) + (size_t)blockIdx.x * 32 * kThreads + threadIdx.x;
#pragma unroll
      for (int n = 0; n < 32; ++n) t[n * kThreads] = x[n];
    }
    const float dt = p.dt[k];
    park_all(C.park0, x);
    // ---- acc = -L F[u] ----
    F.forward(x);
    static_for<0, 4>([&](auto chc) {
      constexpr int ch = decltype(chc)::value;
      float2 v[8];
      spec_chunk<ch>(F, [&](auto ic, int, int, int ft) {
        constexpr int i = decltype(ic)::value;
        const float l = -S.tabL[ft];
        v[i] = make_float2(x[ch * 8 + i].x * l, x[ch * 8 + i].y * l);
      });
      C.park1.store(ch, v);
    });
    C.park1.fence_store();
    __syncthreads();  // P3 reads of this transform finish before the next transform's P1 writes
    // ---- acc += -i kx F[vx u] ----
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
      C.park0.load(ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[ch * 8 + i] = mul2(mul2(v[i], axr), S.ey[4 * (ch * 8 + i) + n2c]);
    }
    F.forward(x);
    static_for<0, 4>([&](auto chc) {
      constexpr int ch = decltype(chc)::value;
      float2 v[8];
      C.park1.load(ch, v);
      spec_chunk<ch>(F, [&](auto ic, int kr, int, int) {
        constexpr int i = decltype(ic)::value;
        const float kk = S.kx[kr];  // (a + ib)(-i kk) = b kk - i a kk
        v[i] = make_float2(fmaf(x[ch * 8 + i].y, kk, v[i].x), fmaf(-x[ch * 8 + i].x, kk, v[i].y));
      });
      C.park1.store(ch, v);
    });
    C.park1.fence_store();
    __syncthreads();
    // ---- acc += -i ky F[vy u];  g = F^-1[m acc] ----
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
      C.park0.load(ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[ch * 8 + i] = mul2(mul2(v[i], exr), S.ay[4 * (ch * 8 + i) + n2c]);
    }
    F.forward(x);
    static_for<0, 4>([&](auto chc) {
      constexpr int ch = decltype(chc)::value;
      float2 v[8];
      C.park1.load(ch, v);
      spec_chunk<ch>(F, [&](auto ic, int, int kc, int ft) {
        constexpr int i = decltype(ic)::value;
        const float kk = S.ky[kc];
        const float m = __fdividef(1.0f / float(kN * kN), fmaf(dt, S.tabA[ft], 1.0f));  // solvers.py:62-63
        const float2 a = make_float2(fmaf(x[ch * 8 + i].y, kk, v[i].x), fmaf(-x[ch * 8 + i].x, kk, v[i].y));
        x[ch * 8 + i] = make_float2(a.x * m, a.y * m);
      });
    });
    F.inverse(x);
    // ---- y1 = u + dt g ----
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
      C.park0.load(ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[ch * 8 + i] = fma2(x[ch * 8 + i], splat2(dt), v[i]);
    }
    __syncthreads();  // all exchange-layout reads done before W is written again
  }
  p1_scatter_nat(F.nb, x);
  __syncthreads();
  ad_store_pair(S.W, p.y1 + oa, p.y1 + ob, C.b_valid);
  ad_teardown(S);
}

// =============================================================================================
// backward (discrete adjoint)
// =============================================================================================
__global__ void __launch_bounds__(kThreads, 1) ad128_bwd_kernel(const __grid_constant__ AdParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  AdSmem& S = *reinterpret_cast<AdSmem*>(smem_raw);
  AdCommon C;
  ad_setup(S, p, C);
  const size_t oa = (size_t)C.env_a * kN * kN, ob = (size_t)C.env_b * kN * kN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  ad_load_pair(S.W, p.y0 + oa, p.y0 + ob);
  __syncthreads();
  const Fft128 F((uint32_t)__cvta_generic_to_shared(S.W), S.tw);
  float2 x[32];
  p1_gather_nat(F.nb, x);
  __syncthreads();
  const int r = F.p1_row();
  const int n2c = F.m1.n2c;

  // gradient accumulators of the current control segment: (cx, cy, p0, p1) for the (a, b) pair
#pragma unroll
  for (int j = 0; j < 4; ++j) S.gacc[j][tid] = make_float2(0.f, 0.f);
  auto flush = [&](int seg) {
    float2 g[4] = {S.gacc[0][tid], S.gacc[1][tid], S.gacc[2][tid], S.gacc[3][tid]};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        g[j].x += __shfl_xor_sync(0xffffffffu, g[j].x, o);
        g[j].y += __shfl_xor_sync(0xffffffffu, g[j].y, o);
      }
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) S.red[warp][j] = g[j];
    }
    __syncthreads();
    if (tid < 8 && p.gctrl != nullptr) {
      const int j = tid & 3, e = tid >> 2;
      float s = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) s += e ? S.red[w][j].y : S.red[w][j].x;
      if (e == 0 || C.b_valid) {
        float* dst = p.gctrl + ((size_t)(e ? C.env_b : C.env_a) * p.nseg + seg) * kAdCtrl + j;
        *dst += s;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) S.gacc[j][tid] = make_float2(0.f, 0.f);
  };

  int cur_seg = -1;
  for (int k = p.ksteps - 1; k >= 0; --k) {
    const int seg = ad_seg(p, k);
    if (seg != cur_seg) {
      if (cur_seg >= 0) flush(cur_seg);
      __syncthreads();
      ad_tables(S, p, C.env_a, C.env_b, seg);
      __syncthreads();
      if (tid == 0) {
        const float* ca = p.ctrl + ((size_t)C.env_a * p.nseg + seg) * kAdCtrl;
        const float* cb = p.ctrl + ((size_t)C.env_b * p.nseg + seg) * kAdCtrl;
        const float2 p0 = make_float2(ca[2], cb[2]), ip1 = make_float2(1.0f / ca[3], 1.0f / cb[3]);
        S.segc[0] = p0;
        S.segc[1] = ip1;
        S.segc[2] = f2mul(p0, ip1);
      }
      __syncthreads();
      cur_seg = seg;
    }
    const float dt = p.dt[k];
    const float2* up = reinterpret_cast<const float2*>(p.traj_in + (size_t)k * p.traj_stride) + (size_t)blockIdx.x * 32 * kThreads + tid;
    // ---- lam_hat = F[lam1] -> park1 ----
    F.forward(x);
    park_all(C.park1, x);
    // ---- d/dx w = F^-1[i kx m lam_hat] ----
    static_for<0, 4>([&](auto chc) {
      constexpr int ch = decltype(chc)::value;
      spec_chunk<ch>(F, [&](auto ic, int kr, int, int ft) {
        constexpr int i = decltype(ic)::value;
        const float m = __fdividef(1.0f / float(kN * kN), fmaf(dt, S.tabA[ft], 1.0f));
        const float km = S.kx[kr] * m;  // (a + ib)(i km) = -b km + i a km
        x[ch * 8 + i] = make_float2(-x[ch * 8 + i].y * km, x[ch * 8 + i].x * km);
      });
    });
    F.inverse(x);
    __syncthreads();
    park_all(C.park0, x);  // d/dx w waits in TMEM until d/dy w is there, so u is read once per step
    // ---- d/dy w = F^-1[i ky m lam_hat] ----
    static_for<0, 4>([&](auto chc) {
      constexpr int ch = decltype(chc)::value;
      float2 v[8];
      C.park1.load(ch, v);
      spec_chunk<ch>(F, [&](auto ic, int, int kc, int ft) {
        constexpr int i = decltype(ic)::value;
        const float m = __fdividef(1.0f / float(kN * kN), fmaf(dt, S.tabA[ft], 1.0f));
        const float km = S.ky[kc] * m;
        x[ch * 8 + i] = make_float2(-v[i].y * km, v[i].x * km);
      });
    });
    F.inverse(x);
    __syncthreads();
    float2 S0 = make_float2(0.f, 0.f), S1 = S0, S2 = S0;
    float2 T0 = S0, T1 = S0, T2 = S0, T3 = S0;
    {
      // vx = vxr * ey, vxr = -(p0/p1) dx ex ;  vy = vyc * dy ey, vyc = -(p0/p1) ex   (row constants
      // re-read from shared memory where used: keeping them in registers across the transforms spilled)
      const float2 coefy = f2scale(f2mul(S.segc[2], S.ex[r]), -dt);
      const float2 coefx = f2mul(coefy, S.dxv[r]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float2 v[8];
        C.park0.load(ch, v);  // d/dx w
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = ch * 8 + i, c = 4 * n + n2c;
          const float2 u = __ldg(up + n * kThreads);
          const float2 px = mul2(v[i], u), py = mul2(x[n], u);
          const float2 ey = S.ey[c], de = S.dyey[c], d2e = S.dy2ey[c];
          S0 = fma2(px, ey, S0);
          S1 = fma2(px, de, S1);
          S2 = fma2(px, d2e, S2);
          T0 = fma2(py, ey, T0);
          T1 = fma2(py, de, T1);
          T2 = fma2(py, d2e, T2);
          T3 = fma2(py, S.dy3ey[c], T3);
          v[i] = fma2(mul2(x[n], coefy), de, mul2(mul2(v[i], coefx), ey));  // dt (vx d/dx w + vy d/dy w)
        }
        C.park0.store(ch, v);
      }
      C.park0.fence_store();
    }
    // ---- lam0 = F^-1[(1 - dt L m) lam_hat] + parked advection part ----
    static_for<0, 4>([&](auto chc) {
      constexpr int ch = decltype(chc)::value;
      float2 v[8];
      C.park1.load(ch, v);
      spec_chunk<ch>(F, [&](auto ic, int, int, int ft) {
        constexpr int i = decltype(ic)::value;
        const float m = __fdividef(1.0f / float(kN * kN), fmaf(dt, S.tabA[ft], 1.0f));
        const float q = fmaf(-dt * S.tabL[ft], m, 1.0f / float(kN * kN));
        x[ch * 8 + i] = make_float2(v[i].x * q, v[i].y * q);
      });
    });
    F.inverse(x);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
      C.park0.load(ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[ch * 8 + i] = add2(x[ch * 8 + i], v[i]);
    }
    __syncthreads();
    // ---- parameter cotangents of this step (per-thread partial sums over its 32 columns) ----
    {
      const float2 dxr = S.dxv[r], exr = S.ex[r], ip1 = S.segc[1], c0 = S.segc[2];
      const float2 vyc = make_float2(-c0.x * exr.x, -c0.y * exr.y), vxr = f2mul(vyc, dxr);
      float2 g_cx = S.gacc[0][tid], g_cy = S.gacc[1][tid], g_p0 = S.gacc[2][tid], g_p1 = S.gacc[3][tid];
      const float2 dx2 = f2mul(dxr, dxr);
      const float2 ip1sq_h = f2scale(f2mul(ip1, ip1), 0.5f);
      const float2 pa = make_float2(fmaf(dx2.x, ip1sq_h.x, -ip1.x), fmaf(dx2.y, ip1sq_h.y, -ip1.y));  // -1/p1 + dx^2/(2 p1^2)
      const float2 dte = f2scale(exr, dt);
      // d vx / d theta
      g_p0 = f2fma(f2mul(f2scale(f2mul(ip1, dxr), -1.0f), dte), S0, g_p0);
      g_cx = f2fma(f2mul(f2mul(c0, dte), make_float2(1.0f - dx2.x * ip1.x, 1.0f - dx2.y * ip1.y)), S0, g_cx);
      const float2 vxd = f2scale(vxr, dt);
      g_cy = f2fma(f2mul(vxd, ip1), S1, g_cy);
      g_p1 = f2fma(f2mul(vxd, pa), S0, g_p1);
      g_p1 = f2fma(f2mul(vxd, ip1sq_h), S2, g_p1);
      // d vy / d theta
      const float2 vyd = f2scale(vyc, dt);
      g_p0 = f2fma(f2mul(f2scale(ip1, -1.0f), dte), T1, g_p0);
      g_cy = f2fma(f2mul(c0, dte), f2sub(T0, f2mul(T2, ip1)), g_cy);
      g_cx = f2fma(f2mul(f2mul(vyd, ip1), dxr), T1, g_cx);
      g_p1 = f2fma(f2mul(vyd, pa), T1, g_p1);
      g_p1 = f2fma(f2mul(vyd, ip1sq_h), T3, g_p1);
      S.gacc[0][tid] = g_cx;
      S.gacc[1][tid] = g_cy;
      S.gacc[2][tid] = g_p0;
      S.gacc[3][tid] = g_p1;
    }
  }
  if (cur_seg >= 0) flush(cur_seg);
  __syncthreads();
  p1_scatter_nat(F.nb, x);
  __syncthreads();
  ad_store_pair(S.W, p.y1 + oa, p.y1 + ob, C.b_valid);
  ad_teardown(S);
}

}  // namespace pdeopt
