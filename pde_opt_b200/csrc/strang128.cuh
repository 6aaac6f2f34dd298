// Fused K-step Strang split-step stepper for 128x128 complex fields (sm_100a).
//
// Replaces K iterations of the diffeqsolve loop body of the reference with
//   StrangSplitting.step                 pde_opt/numerics/solvers.py:99-122
//   GPE2DTSControl.B_terms               pde_opt/numerics/equations/gross_pitaevskii.py:67-75
// for one environment per CTA: the wavefunction [128][128][2] (re, im) is exactly the float2
// pair layout of the FFT core (fft128.cuh), so no packing is needed.
//
// Per step (reference line numbers in solvers.py):
//   dt_c = (t1 - t0) * time_scale                       :101
//   tmp  = ifft( fft(psi0) * exp(A_term * dt_c / 2) )   :105-108   (skipped when A_term == 0, F8)
//   tmp *= exp( b(psi0) * dt_c ),  b = -i V,            :109-110   b evaluated at psi0 (quirk kept)
//          V = trap/2 ((1+e) x^2 + (1-e) y^2) + lights(x, y) + k |psi0|^2
//   tmp /= sqrt( sum |tmp|^2 dx^2 )                     :111       CTA-wide reduction
//   y1   = ifft( fft(tmp) * exp(A_term * dt_c / 2) )    :112-114
// psi0 is parked in TMEM while the registers hold FFT data.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft128.cuh"
#include "sifs128.cuh"

namespace pdeopt {

struct StrangParams {
  const float* y0;      // [batch][128][128][2]
  float* y1;
  int batch, ksteps;
  const float* a_term;  // folded [65*65][2] complex A_term, or null when it is identically zero
  float ts_re, ts_im;   // time_scale (solvers.py:90)
  float dx;             // solvers.py:111
  float k_int, e, trap; // gross_pitaevskii.py:38-44
  float lo_x, lo_y, hx, hy;
  const float* ctrl;    // [batch][kNCtrl]: [1] amp [2] x0 [3] y0 [4] width of a Gaussian `lights` spot, or null
  float* park;
  float dt[kMaxK];
};

struct __align__(1024) StrangSmem {
  float2 W[kN * kN];
  float2 atab[kTabLen + 1];
  float2 tw[128];
  float gx[kN], gy[kN];
  float red[kThreads / 32];
  uint32_t tmem_base;
};

__device__ __forceinline__ float2 cexp_times(float2 a, float hr, float hi, float scale) {
  // scale * exp((a.x + i a.y) * (hr + i hi))
  const float re = a.x * hr - a.y * hi, im = a.x * hi + a.y * hr;
  const float m = __expf(re) * scale;
  float s, c;
  const float red = im - 6.283185307179586f * rintf(im * 0.15915494309189535f);
  __sincosf(red, &s, &c);
  return make_float2(m * c, m * s);
}

__global__ void __launch_bounds__(kThreads, 1) strang128_kernel(const __grid_constant__ StrangParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  StrangSmem& S = *reinterpret_cast<StrangSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int env = blockIdx.x;
  const bool has_A = p.a_term != nullptr;

  Park park;
#ifndef PDEOPT_PARK_GLOBAL
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&S.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
#endif
  if (tid < 128) {
    float s, c;
    sincospif(-2.0f * float(tid) / 128.0f, &s, &c);
    S.tw[tid] = make_float2(c, s);
  }
  bool has_light = false;
  if (p.ctrl != nullptr) {
    const float* cc = p.ctrl + (size_t)env * kNCtrl;
    has_light = cc[1] != 0.f;
    if (tid < 2 * kN) {
      const int i = tid & (kN - 1);
      const bool isx = tid < kN;
      const float pos = isx ? (p.lo_x + (i + 0.5f) * p.hx) : (p.lo_y + (i + 0.5f) * p.hy);
      const float d = pos - (isx ? cc[2] : cc[3]);
      const float v = has_light ? expf(-d * d * 0.5f / (cc[4] * cc[4])) * (isx ? cc[1] : 1.0f) : 0.f;
      if (isx) S.gx[i] = v; else S.gy[i] = v;
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

  // ---- prologue: psi0 -> natural layout -> P1 registers ----
  {
    const float2* src = reinterpret_cast<const float2*>(p.y0) + (size_t)env * kN * kN;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      const float4 a = *reinterpret_cast<const float4*>(src + r * kN + 4 * lane);
      const float4 b = *reinterpret_cast<const float4*>(src + r * kN + 4 * lane + 2);
      float2 v[4] = {make_float2(a.x, a.y), make_float2(a.z, a.w), make_float2(b.x, b.y), make_float2(b.z, b.w)};
      store_row(S.W, r, lane, v);
    }
  }
  __syncthreads();
  const Fft128 F((uint32_t)__cvta_generic_to_shared(S.W), S.tw);
  float2 x[32];
  p1_gather_nat(F.nb, x);
  __syncthreads();  // natural layout fully read before the buffer becomes exchange space

  // position-dependent part of V in the P1 arrangement: row term (one value) and column terms
  const int r = F.p1_row();
  const float xr = p.lo_x + (r + 0.5f) * p.hx;
  const float vrow = 0.5f * p.trap * (1.0f + p.e) * xr * xr;
  const float gxr = has_light ? S.gx[r] : 0.f;

  // exp(A_term * 0.5 * dt_c) / N^2 is tabulated in shared memory per distinct dt (the step lengths of
  // the float32 time grid differ by ulps, so the table is rebuilt only when dt changes): the half
  // kinetic step is then one LDS.64 + one complex multiply per element instead of exp + sincos.
  float dt_tab = __int_as_float(0x7fc00000);
  auto build_etab = [&](float dt) {
    const float hr = 0.5f * dt * p.ts_re, hi = 0.5f * dt * p.ts_im;
    __syncthreads();  // earlier readers of the table are done
    for (int i = tid; i < kTabLen; i += kThreads)
      S.atab[i] = cexp_times(reinterpret_cast<const float2*>(p.a_term)[i], hr, hi, 1.0f / float(kN * kN));
    __syncthreads();
  };
  auto half_kinetic = [&]() {
    static_for<0, 2>([&](auto bc) {
      constexpr int b = decltype(bc)::value;
      const int kc = F.p3_kc(b);
      const int fc = kc <= 64 ? kc : 128 - kc;
      static_for<0, 16>([&](auto pc) {
        constexpr int pp = decltype(pc)::value;
        const int kr = F.p3_kr(pp);
        const int fr = kr <= 64 ? kr : 128 - kr;
        x[b * 16 + pp] = cmul(x[b * 16 + pp], S.atab[fr * kTabDim + fc]);
      });
    });
  };

  for (int k = 0; k < p.ksteps; ++k) {
    const float dt = p.dt[k];
    if (has_A && dt != dt_tab) {
      build_etab(dt);
      dt_tab = dt;
    }
    // park psi0 (needed for b(psi0) after the first half step)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = x[ch * 8 + i];
      park.store(ch, v);
    }
    park.fence_store();
    if (has_A) {
      F.forward(x);
      half_kinetic();
      F.inverse(x);
    }
    // tmp *= exp(b(psi0) dt_c);  b dt_c = V dt (ts_im - i ts_re)
    float part = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
      park.load(ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = ch * 8 + i;
        const int c = F.p1_col(n);
        const float yc = p.lo_y + (c + 0.5f) * p.hy;
        float V = vrow + 0.5f * p.trap * (1.0f - p.e) * yc * yc + p.k_int * (v[i].x * v[i].x + v[i].y * v[i].y);
        if (has_light) V = fmaf(gxr, S.gy[c], V);
        const float a = V * dt;
        if (p.ts_re == 0.f) {  // imaginary time: real factor (uniform branch)
          const float m = __expf(a * p.ts_im);
          x[n] = make_float2(x[n].x * m, x[n].y * m);
        } else {
          const float ph = -a * p.ts_re;
          float s, cth;
          __sincosf(ph - 6.283185307179586f * rintf(ph * 0.15915494309189535f), &s, &cth);
          if (p.ts_im != 0.f) {
            const float m = __expf(a * p.ts_im);
            s *= m;
            cth *= m;
          }
          x[n] = cmul(x[n], make_float2(cth, s));
        }
        part = fmaf(x[n].x, x[n].x, fmaf(x[n].y, x[n].y, part));
      }
    }
    // global renormalisation (solvers.py:111)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) S.red[warp] = part;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) tot += S.red[w];
    __syncthreads();  // S.red is rewritten next step
    const float scale = rsqrtf(tot * p.dx * p.dx);
#pragma unroll
    for (int n = 0; n < 32; ++n) x[n] = make_float2(x[n].x * scale, x[n].y * scale);
    if (has_A) {
      F.forward(x);
      half_kinetic();
      F.inverse(x);
    }
  }

  // ---- epilogue ----
  __syncthreads();
  p1_scatter_nat(F.nb, x);
  __syncthreads();
  {
    float2* dst = reinterpret_cast<float2*>(p.y1) + (size_t)env * kN * kN;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = warp * 8 + i;
      float2 v[4];
      load_row(S.W, rr, lane, v);
      *reinterpret_cast<float4*>(dst + rr * kN + 4 * lane) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
      *reinterpret_cast<float4*>(dst + rr * kN + 4 * lane + 2) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
    }
  }
#ifndef PDEOPT_PARK_GLOBAL
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(S.tmem_base));
#endif
}

}  // namespace pdeopt
