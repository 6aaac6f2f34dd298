// Fused K-step semi-implicit stepper with the pseudo-spectral right-hand sides (derivs="fourier")
// for 128x128 fields (sm_100a).
//
// Replaces, per step,
//   CahnHilliard2DPeriodic.rhs_fourier   pde_opt/numerics/equations/cahn_hilliard.py:82-87
//       tmp = fft(mu(u)) - kappa k2 fft(u);  rhs = Re ifft( sum_q q fft( D(u) ifft(q tmp) ) ),  q = 2 pi i k_{x,y}
//   AllenCahn2DPeriodic.rhs_fourier      pde_opt/numerics/equations/allen_cahn.py:74-79
//       mu = ifft( fft(mu(u)) - kappa k2 fft(u) );  rhs = -R(u) Re mu
//   SemiImplicitFourierSpectral.step     pde_opt/numerics/solvers.py:56-70
// The reference carries COMPLEX intermediates (the odd multipliers 2 pi i k are not zeroed on the
// Nyquist lines, so ifft(q tmp) has a small imaginary part that D(u) then spreads over all modes);
// to reproduce that exactly this kernel keeps one environment per CTA in full complex arithmetic
// instead of the two-environments-per-complex-field packing of the finite-difference kernel:
// 9 complex FFTs per Cahn-Hilliard step (5 for Allen-Cahn).
//
// On-chip: registers = FFT working set, TMEM slot 0 = state u, slot 1 = the spectrum `tmp`; the
// x-part of the flux divergence waits in a per-CTA global scratch line (L2-resident, coalesced).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "spec_util.cuh"
#include "pointwise.cuh"
#include "sifs128.cuh"

namespace pdeopt {

struct FourierParams {
  const float* y0;
  float* y1;
  int batch, ksteps, mode, eq;  // mode: MODE_FUSED / MODE_RHS_ONLY ; eq: EQ_CH / EQ_AC
  const float* symbol;          // [65*65] folded A*sigma (null in rhs-only mode)
  const float* kx;              // [128] imag(two_pi_i_kx) along axis 0 (NOT zeroed at Nyquist)
  const float* ky;              // [128]
  const float* ctrl;            // [batch][8] or null
  float2* scratch;              // [grid][32][512] complex
  float kappa, lo_x, lo_y, hx, hy;
  PointwiseParams pw;
  float dt[kMaxK];
};

struct __align__(1024) FourierSmem {
  float2 W[kN * kN];
  float tab[kTabLen + 3];
  float kx[kN], ky[kN];
  float2 tw[128];
  float gx[kN], gy[kN];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1) fourier128_kernel(const __grid_constant__ FourierParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FourierSmem& S = *reinterpret_cast<FourierSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int env = blockIdx.x;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&S.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (p.symbol != nullptr)
    for (int i = tid; i < kTabLen; i += kThreads) S.tab[i] = p.symbol[i];
  float w_off = 0.f;
  bool has_bump = false;
  if (p.ctrl != nullptr) {
    const float* cc = p.ctrl + (size_t)env * kNCtrl;
    w_off = cc[0];
    has_bump = cc[1] != 0.f;
  }
  if (tid < kN) {
    S.kx[tid] = p.kx[tid];
    S.ky[tid] = p.ky[tid];
    float s, c;
    sincospif(-2.0f * float(tid) / 128.0f, &s, &c);
    S.tw[tid] = make_float2(c, s);
  }
  if (has_bump && tid < 2 * kN) {
    const float* cc = p.ctrl + (size_t)env * kNCtrl;
    const int i = tid & (kN - 1);
    const bool isx = tid < kN;
    const float pos = isx ? (p.lo_x + (i + 0.5f) * p.hx) : (p.lo_y + (i + 0.5f) * p.hy);
    const float d = pos - (isx ? cc[2] : cc[3]);
    const float v = expf(-d * d * 0.5f / (cc[4] * cc[4])) * (isx ? cc[1] : 1.0f);
    if (isx) S.gx[i] = v; else S.gy[i] = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  Park park0, park1;
  park0.taddr = S.tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;
  park1.taddr = park0.taddr + 256u;
  float2* G = p.scratch + (size_t)blockIdx.x * 32 * kThreads;

  // ---- prologue: u -> natural layout as (u, 0) -> P1 registers ----
  {
    const float* ya = p.y0 + (size_t)env * kN * kN;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      const float4 a = *reinterpret_cast<const float4*>(ya + r * kN + 4 * lane);
      float2 v[4] = {make_float2(a.x, 0.f), make_float2(a.y, 0.f), make_float2(a.z, 0.f), make_float2(a.w, 0.f)};
      store_row(S.W, r, lane, v);
    }
  }
  __syncthreads();
  const Fft128 F((uint32_t)__cvta_generic_to_shared(S.W), S.tw);
  float2 x[32];
  p1_gather_nat(F.nb, x);
  __syncthreads();
  const int r = F.p1_row(), n2c = F.m1.n2c;
  const float gxr = has_bump ? S.gx[r] : 0.f;
  const float inv_n2 = 1.0f / float(kN * kN);

  const int nsteps = (p.mode == MODE_RHS_ONLY) ? 1 : p.ksteps;
  for (int k = 0; k < nsteps; ++k) {
    park_all(park0, x);  // u
    // ---- A: -kappa k2 fft(u)  (k2 = two_pi_i_k_2 = -(KX^2 + KY^2)) ----
    F.forward(x);
    static_for<0, 4>([&](auto chc) {
      constexpr int ch = decltype(chc)::value;
      float2 v[8];
      spec_chunk<ch>(F, [&](auto ic, int kr, int kc, int) {
        constexpr int i = decltype(ic)::value;
        const float k2 = -(S.kx[kr] * S.kx[kr] + S.ky[kc] * S.ky[kc]);
        const float l = -(p.kappa * k2);
        v[i] = make_float2(x[ch * 8 + i].x * l, x[ch * 8 + i].y * l);
      });
      park1.store(ch, v);
    });
    park1.fence_store();
    __syncthreads();
    // ---- B: tmp = fft(mu_h(u)) - kappa k2 fft(u) ----
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
      park0.load(ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float m = mu_h<MU_RUNTIME>(v[i].x, p.pw, w_off);
        if (has_bump) m = fmaf(gxr, S.gy[4 * (ch * 8 + i) + n2c], m);
        x[ch * 8 + i] = make_float2(m, 0.f);
      }
    }
    F.forward(x);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
      park1.load(ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = f2add(x[ch * 8 + i], v[i]);
        x[ch * 8 + i] = v[i];
      }
      if (p.eq == EQ_CH) park1.store(ch, v);
    }
    park1.fence_store();
    if (p.eq == EQ_AC) {
      // mu = ifft(tmp); f = -R(u) Re mu   (allen_cahn.py:77-79)
#pragma unroll
      for (int n = 0; n < 32; ++n) x[n] = f2scale(x[n], inv_n2);
      F.inverse(x);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float2 v[8];
        park0.load(ch, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[ch * 8 + i] = make_float2(-mob<MOB_RUNTIME>(v[i].x, p.pw) * x[ch * 8 + i].x, 0.f);
      }
    } else {
      // ---- C: acc_x = q_x fft( D(u) ifft(q_x tmp) ),  q_x = i KX ----
      static_for<0, 4>([&](auto chc) {
        constexpr int ch = decltype(chc)::value;
        spec_chunk<ch>(F, [&](auto ic, int kr, int, int) {
          constexpr int i = decltype(ic)::value;
          const float q = S.kx[kr] * inv_n2;  // (a + ib)(i q) = -b q + i a q
          x[ch * 8 + i] = make_float2(-x[ch * 8 + i].y * q, x[ch * 8 + i].x * q);
        });
      });
      F.inverse(x);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float2 v[8];
        park0.load(ch, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[ch * 8 + i] = f2scale(x[ch * 8 + i], mob<MOB_RUNTIME>(v[i].x, p.pw));
      }
      F.forward(x);
      static_for<0, 4>([&](auto chc) {
        constexpr int ch = decltype(chc)::value;
        spec_chunk<ch>(F, [&](auto ic, int kr, int, int) {
          constexpr int i = decltype(ic)::value;
          const float q = S.kx[kr];
          G[(ch * 8 + i) * kThreads + tid] = make_float2(-x[ch * 8 + i].y * q, x[ch * 8 + i].x * q);
        });
      });
      // ---- D: acc = acc_x + q_y fft( D(u) ifft(q_y tmp) ) ----
      static_for<0, 4>([&](auto chc) {
        constexpr int ch = decltype(chc)::value;
        float2 v[8];
        park1.load(ch, v);
        spec_chunk<ch>(F, [&](auto ic, int, int kc, int) {
          constexpr int i = decltype(ic)::value;
          const float q = S.ky[kc] * inv_n2;
          x[ch * 8 + i] = make_float2(-v[i].y * q, v[i].x * q);
        });
      });
      F.inverse(x);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float2 v[8];
        park0.load(ch, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[ch * 8 + i] = f2scale(x[ch * 8 + i], mob<MOB_RUNTIME>(v[i].x, p.pw));
      }
      F.forward(x);
      static_for<0, 4>([&](auto chc) {
        constexpr int ch = decltype(chc)::value;
        spec_chunk<ch>(F, [&](auto ic, int, int kc, int) {
          constexpr int i = decltype(ic)::value;
          const float q = S.ky[kc];
          const float2 g = G[(ch * 8 + i) * kThreads + tid];
          x[ch * 8 + i] = make_float2(fmaf(-x[ch * 8 + i].y, q, g.x) * inv_n2, fmaf(x[ch * 8 + i].x, q, g.y) * inv_n2);
        });
      });
      // ---- E: f0 = Re ifft(acc) ----
      F.inverse(x);
#pragma unroll
      for (int n = 0; n < 32; ++n) x[n].y = 0.f;
    }
    if (p.mode == MODE_RHS_ONLY) break;
    // ---- F: y1 = u + dt Re ifft( fft(f0) / (1 + A dt sigma) )   (solvers.py:62-63) ----
    const float dt = p.dt[k];
    F.forward(x);
    static_for<0, 4>([&](auto chc) {
      constexpr int ch = decltype(chc)::value;
      spec_chunk<ch>(F, [&](auto ic, int, int, int ft) {
        constexpr int i = decltype(ic)::value;
        const float m = __fdividef(inv_n2, fmaf(dt, S.tab[ft], 1.0f));
        x[ch * 8 + i] = make_float2(x[ch * 8 + i].x * m, x[ch * 8 + i].y * m);
      });
    });
    F.inverse(x);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
      park0.load(ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[ch * 8 + i] = make_float2(fmaf(dt, x[ch * 8 + i].x, v[i].x), 0.f);
    }
  }
  // ---- epilogue ----
  __syncthreads();
  p1_scatter_nat(F.nb, x);
  __syncthreads();
  {
    float* ya = p.y1 + (size_t)env * kN * kN;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = warp * 8 + i;
      float2 v[4];
      load_row(S.W, rr, lane, v);
      *reinterpret_cast<float4*>(ya + rr * kN + 4 * lane) = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(S.tmem_base));
}

}  // namespace pdeopt
