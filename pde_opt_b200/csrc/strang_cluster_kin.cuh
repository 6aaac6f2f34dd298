// Fused K-step Strang split step of GPE2DTSControl WITH the kinetic term on 256x256 complex64 fields:
// one thread-block CLUSTER of 4 CTAs x 512 threads per environment (8 x 256 with PDEOPT_CF_CTAS=8, see cfft256.cuh), the wavefunction resident in the cluster's
// registers / shared memory / tensor memory for all K steps (sm_100a).
//
// Replaces K calls of StrangSplitting.step (pde_opt/numerics/solvers.py:99-122) with
// GPE2DTSControl.B_terms (gross_pitaevskii.py:67-75) evaluated at y0:
//   tmp = ifft2( fft2(psi0) * exp(A dt_c / 2) )
//   tmp *= exp(b(psi0) dt_c);  tmp /= sqrt(sum |tmp|^2 dx^2)
//   y1  = ifft2( fft2(tmp) * exp(A dt_c / 2) )
// Four 2-D transforms per step = eight line passes over 64 local lines per CTA (cfft256.cuh); the four
// row <-> column transposes per step cross the cluster as st.shared::cluster stores straight into the
// slot where the next pass reads (no pack / unpack pass, no HBM or L2 traffic for the state); the norm
// is one float per CTA through the same channel and the renormalisation rides on the second kinetic
// multiplier.  psi0 of the step is parked in 256 TMEM columns.  HBM traffic: the state once in and once
// out per launch, plus the 512 KB kinetic table per distinct dt (L2-resident, shared by all clusters).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cfft256.cuh"
#include "sifs128.cuh"          // tmem_ld16 / tmem_st16, kMaxK, kNCtrl
#include "strang_cluster.cuh"   // cluster_ctarank, cluster_sync_all, st_cluster_f32

namespace pdeopt {
namespace cf {

constexpr int kMaxTabs = 8;  // distinct dt per launch

struct KinParams {
  const float* y0;  // [batch][256][256][2]
  float* y1;
  int batch, ksteps;
  float ts_re, ts_im, dx;
  float k_int, e, trap;
  float lo_x, lo_y, hx, hy;
  const float* ctrl;      // [batch][8] or null (Gaussian light spot)
  const float2* etab;     // [n_tabs][256 kc][256 kr]: exp(A_term[kr][kc] dt_c / 2) / 65536, one table per distinct dt
  float dt[kMaxK];
  unsigned char tab[kMaxK];  // table index of every step
};

struct __align__(2048) KinSmem {
  unsigned char slab[kSlabBytes];
  float2 tw[8 * 32];
  float gx[kLines], gy[kN];
  float red[kThreadsC / 32];
  float part[2][kCtas];
  uint32_t tp[kThreadsC * kCtas];  // per-thread transposed-store bases (cfft256.cuh LineMap)
  uint32_t tmem_base;
};

// exp(A dt_c / 2) / N^2 of one dt, transposed ([kc][kr]) so that the column pass reads it along kr
__global__ void strang_cluster_etab_kernel(const float2* __restrict__ a_term, float2* __restrict__ etab, float hr, float hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kN * kN) return;
  const int kc = i / kN, kr = i % kN;
  const float2 a = a_term[(size_t)kr * kN + kc];
  const float re = a.x * hr - a.y * hi, im = a.x * hi + a.y * hr;
  const float m = expf(re) / float(kN * kN);
  float s, c;
  sincosf(im, &s, &c);
  etab[i] = make_float2(m * c, m * s);
}

// one kinetic half step: rows forward, transpose, columns forward * table * inverse, transpose, rows inverse.
// x: spatial arrangement of row gl in (natural n1 order) and out.  `part`: the four partial norms of the
// preceding potential step (with_norm), visible after the first cluster barrier below.
__device__ __forceinline__ void kinetic_half(const Ctx& c, const LineMap& m, const float2* __restrict__ tw, const float* part,
                                             const float2* __restrict__ trow, float dx, bool with_norm, float2 (&x)[32]) {
  {
    float2 y[32];
    static_for<0, 32>([&](auto nc) { y[brev<5>(decltype(nc)::value)] = x[decltype(nc)::value]; });
    line_fwd(c, tw, m, y);
    cluster_sync_exec();  // A: every CTA is done with its slab (loads and in-line exchanges)
    store_transposed_from_freq(c, m, y);
  }
  cluster_sync_all();  // B: the column slabs have arrived
  float scale = 1.0f;
  if (with_norm) {
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < kCtas; ++r) tot += part[r];
    scale = rsqrtf(tot * dx * dx);  // solvers.py:111, applied with the multiplier (the transforms are linear)
  }
  load_spatial(c, m, x);
  line_fwd(c, tw, m, x);
  {
    // trow = table + 256 gl + j: kc = gl, kr = j + 8 a + 32 k0
    // in groups of eight: all 32 table loads in flight at once would need 64 more registers next to the 64 of x
    static_for<0, 4>([&](auto gc) {
      constexpr int g0 = 8 * decltype(gc)::value;
      float2 mm[8];
      static_for<0, 8>([&](auto ic) {
        constexpr int i = g0 + decltype(ic)::value;
        mm[i - g0] = __ldg(trow + 8 * (i >> 3) + 32 * (i & 7));
      });
      static_for<0, 8>([&](auto ic) {
        constexpr int i = g0 + decltype(ic)::value;
        x[i] = cmul(x[i], make_float2(mm[i - g0].x * scale, mm[i - g0].y * scale));
      });
      asm volatile("" ::: "memory");
    });
  }
  line_inv(c, tw, m, x);
  cluster_sync_exec();  // A
  store_transposed_from_spatial(c, m, x);
  cluster_sync_all();  // B: the row slabs (frequency along the row) have arrived
  load_freq(c, m, x);
  line_inv(c, tw, m, x);
}

static __global__ void __launch_bounds__(kThreadsC, kCtas == 8 ? 2 : 1) strang_cluster_kin_kernel(const __grid_constant__ KinParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the slab must be 2048-byte aligned in the shared window (the swizzle XORs of cfft256.cuh commute with the base)
  const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  KinSmem& S = *reinterpret_cast<KinSmem*>(smem_raw + (((raw_s + 2047u) & ~2047u) - raw_s));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l = thread_line(tid), j = thread_j(tid);
  const uint32_t q = cluster_ctarank();
  const int env = blockIdx.x / kCtas;
  const int gl = (int)q * kLines + l;  // global line index of this thread's line in either slab orientation

  if (warp == 0) {
    // psi0 of the step: 64 columns per thread, two warps per lane quadrant
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&S.tmem_base)), "n"(kThreadsC / 2));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid < 256) {  // kThreadsC >= 256
    float s, c;
    sincospif(-2.0f * float((tid >> 5) * (tid & 31)) / 256.0f, &s, &c);
    S.tw[tid] = make_float2(c, s);
  }
  bool has_light = false;
  if (p.ctrl != nullptr) {
    const float* cc = p.ctrl + (size_t)env * kNCtrl;
    has_light = cc[1] != 0.f;
    if (has_light) {
      if (tid < kN) {
        const float d = p.lo_y + (tid + 0.5f) * p.hy - cc[3];
        S.gy[tid] = expf(-d * d * 0.5f / (cc[4] * cc[4]));
      }
      if (tid < kLines) {
        const float d = p.lo_x + ((int)q * kLines + tid + 0.5f) * p.hx - cc[2];
        S.gx[tid] = cc[1] * expf(-d * d * 0.5f / (cc[4] * cc[4]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = S.tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;

  Ctx c;
  c.base = (uint32_t)__cvta_generic_to_shared(S.slab);
#pragma unroll
  for (int r = 0; r < kCtas; ++r) c.peer[r] = (r == (int)q) ? c.base : peer_addr(c.base, (uint32_t)r);
  c.self = (int)q;
  const uint32_t part_saddr = (uint32_t)__cvta_generic_to_shared(&S.part[0][0]);
  const LineMap lm(c, l, j, gl, (uint32_t)__cvta_generic_to_shared(S.tp), tid);

  // ---- load: coalesced global reads of this CTA's 64 rows into the slab, then the spatial arrangement
  // (thread (l, j) holds columns 8 n1 + j of row gl) ----
  float2 x[32];
  {
    const float2* src = reinterpret_cast<const float2*>(p.y0) + ((size_t)env * kN + (size_t)q * kLines) * kN;
#pragma unroll 8
    for (int it = 0; it < 32; ++it) {
      const int i = it * kThreadsC + tid;
      c.st<0>(c.base + slot_bytes(i >> 8, i & 255), __ldg(src + i));
    }
    __syncthreads();
    float2 y[32];
    load_spatial(c, lm, y);
    static_for<0, 32>([&](auto nc) { x[decltype(nc)::value] = y[brev<5>(decltype(nc)::value)]; });
  }
  const float xr = p.lo_x + (gl + 0.5f) * p.hx;
  const float vrow = 0.5f * p.trap * (1.0f + p.e) * xr * xr;
  const float vcy = 0.5f * p.trap * (1.0f - p.e);
  const float gxl = has_light ? S.gx[l] : 0.f;
  const bool pure_imag = p.ts_re == 0.f, pure_real = p.ts_im == 0.f;  // uniform
  cluster_sync_all();  // every CTA of the cluster is running and its shared memory is valid

  for (int k = 0; k < p.ksteps; ++k) {
    const float dt = p.dt[k];
    const float2* tab = p.etab + (size_t)p.tab[k] * kN * kN;
    // opaque copy of j: everything derived from it (twiddle pointer, column coordinates) is recomputed inside
    // the step instead of being hoisted out of the loop into 100+ registers' worth of local-memory spills
    // (a cluster barrier invalidates L1, so every spilled value would come back from L2)
    int jv = j;
    asm volatile("" : "+r"(jv));
    const float2* twj = S.tw + 32 * jv;
    // psi0 of the step -> tensor memory (b is evaluated at y0, solvers.py:109)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = x[ch * 8 + i];
      tmem_st16(taddr + ch * 16, v);
    }
    tmem_wait_st();
    kinetic_half(c, lm, twj, S.part[0], tab + (size_t)gl * kN + j, p.dx, false, x);
    // potential step with b(psi0), partial norm
    float acc = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float2 p0[8];
      tmem_ld16(taddr + ch * 16, p0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n1 = ch * 8 + i;
        const int col = 8 * n1 + jv;
        const float yc = p.lo_y + (col + 0.5f) * p.hy;
        float V = fmaf(vcy * yc, yc, vrow) + p.k_int * (p0[i].x * p0[i].x + p0[i].y * p0[i].y);
        if (has_light) V = fmaf(gxl, S.gy[col], V);
        const float a = V * dt;
        float2 w;
        if (pure_imag) {
          const float m = __expf(a * p.ts_im);
          w = make_float2(x[n1].x * m, x[n1].y * m);
        } else {
          const float ph = -a * p.ts_re;
          float s, cth;
          __sincosf(ph - 6.283185307179586f * rintf(ph * 0.15915494309189535f), &s, &cth);
          if (!pure_real) {
            const float m = __expf(a * p.ts_im);
            s *= m;
            cth *= m;
          }
          w = cmul(x[n1], make_float2(cth, s));
        }
        x[n1] = w;
        acc = fmaf(w.x, w.x, fmaf(w.y, w.y, acc));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) S.red[warp] = acc;
    __syncthreads();
    if (tid < kCtas) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kThreadsC / 32; ++w) t += S.red[w];
      st_cluster_f32(part_saddr + (uint32_t)(((k & 1) * kCtas + (int)q) * sizeof(float)), (uint32_t)tid, t);
    }
    // (the partial sums become visible at the cluster barriers inside the second kinetic half step,
    // before its multiplier is applied; `red` is not touched again before the next step's __syncthreads)
    kinetic_half(c, lm, twj, S.part[k & 1], tab + (size_t)gl * kN + j, p.dx, true, x);
  }
  {
    // through the slab again so that the global stores are coalesced
    __syncthreads();
    store_spatial(c, lm, x);
    __syncthreads();
    float2* dst = reinterpret_cast<float2*>(p.y1) + ((size_t)env * kN + (size_t)q * kLines) * kN;
#pragma unroll 8
    for (int it = 0; it < 32; ++it) {
      const int i = it * kThreadsC + tid;
      dst[i] = c.ld<0>(c.base + slot_bytes(i >> 8, i & 255));
    }
  }
  cluster_sync_all();  // no CTA exits while a peer may still write into its shared memory
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(S.tmem_base), "n"(kThreadsC / 2));
  }
}

}  // namespace cf
}  // namespace pdeopt
