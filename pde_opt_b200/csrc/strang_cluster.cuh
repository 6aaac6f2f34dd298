// Fused K-step Strang step of GPE2DTSControl AS SHIPPED (A_term == 0, gross_pitaevskii.py:62) on
// 256x256 complex64 fields: one thread-block CLUSTER of 4 CTAs per environment (sm_100a).
//
// With the kinetic term disabled the step (solvers.py:99-122) is pointwise,
//     psi <- psi * exp(b(psi) dt_c) / || psi * exp(b(psi) dt_c) ||,
// except for the norm, a reduction over the whole field.  The 512 KB wavefunction does not fit one
// SM, so four CTAs hold 128 KB each IN REGISTERS (32 complex values per thread) for all K steps and
// exchange only one partial sum per step through distributed shared memory (st.shared::cluster +
// barrier.cluster): the state is read from HBM once and written once per launch, instead of once per
// step as on the line-FFT path.  (With a non-zero A_term the transposes between row and column
// transforms would have to cross DSMEM at ~20 B/clk/SM; that case stays on the L2-resident line path.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sifs128.cuh"

namespace pdeopt {

constexpr int kClN = 256;       // grid size per axis
constexpr int kClCtas = 4;      // CTAs per cluster (64 rows each)
constexpr int kClThreads = 512;

struct StrangClusterParams {
  const float* y0;  // [batch][256][256][2]
  float* y1;
  int batch, ksteps;
  float ts_re, ts_im, dx;
  float k_int, e, trap;
  float lo_x, lo_y, hx, hy;
  const float* ctrl;  // [batch][8] or null (Gaussian light spot)
  float dt[kMaxK];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution-only cluster barrier: orders a write-after-read hazard (every thread of every CTA has FINISHED READING its
// shared memory — the loaded values were consumed before the arrive — so peers may overwrite it).  No release fence: this CTA
// has no remote stores in flight at such a point, and the release variant stalls on a membar anyway (10 % of the
// kinetic kernel's stall samples).
__device__ __forceinline__ void cluster_sync_exec() {
#ifdef PDEOPT_CLUSTER_SYNC_RELEASE_ONLY  // A/B switch: release + acquire everywhere (616 k vs 632 k env-steps/s at 128 environments)
  cluster_sync_all();
#else
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void st_cluster_f32(uint32_t local_saddr, uint32_t rank, float v) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_saddr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

static __global__ void __launch_bounds__(kClThreads, 1) strang_cluster_kernel(const __grid_constant__ StrangClusterParams p) {
  __shared__ float red[kClThreads / 32];
  __shared__ float part[2][kClCtas];  // per-step partial sums of the 4 CTAs, double-buffered by step parity
  __shared__ float gx[kClN / kClCtas], gy[kClN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t q = cluster_ctarank();
  const int env = blockIdx.x / kClCtas;
  const int col = tid & (kClN - 1), rsub = tid >> 8;  // element j: local row 2 j + rsub, column col
  const int row0 = (int)q * (kClN / kClCtas);

  bool has_light = false;
  if (p.ctrl != nullptr) {
    const float* cc = p.ctrl + (size_t)env * kNCtrl;
    has_light = cc[1] != 0.f;
    if (has_light) {
      if (tid < kClN) {
        const float d = p.lo_y + (tid + 0.5f) * p.hy - cc[3];
        gy[tid] = expf(-d * d * 0.5f / (cc[4] * cc[4]));
      }
      if (tid < kClN / kClCtas) {
        const float d = p.lo_x + (row0 + tid + 0.5f) * p.hx - cc[2];
        gx[tid] = cc[1] * expf(-d * d * 0.5f / (cc[4] * cc[4]));
      }
    }
  }
  // ---- load: element j of thread t is float2 index j*512 + t of this CTA's 64-row block (coalesced)
  const float2* src = reinterpret_cast<const float2*>(p.y0) + (size_t)env * kClN * kClN + (size_t)row0 * kClN;
  float2 x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = src[j * kClThreads + tid];
  const float yc = p.lo_y + (col + 0.5f) * p.hy;
  const float vcol = 0.5f * p.trap * (1.0f - p.e) * yc * yc;
  const float gyc = has_light ? gy[col] : 0.f;  // visible after the barrier below
  cluster_sync_all();                          // all CTAs of the cluster are running; tables written
  const float gyv = has_light ? gy[col] : gyc;
  const uint32_t part_saddr = (uint32_t)__cvta_generic_to_shared(&part[0][0]);

  const bool pure_imag = p.ts_re == 0.f, pure_real = p.ts_im == 0.f;  // uniform
  for (int k = 0; k < p.ksteps; ++k) {
    const float dt = p.dt[k];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int rl = 2 * j + rsub;
      const float xr = p.lo_x + (row0 + rl + 0.5f) * p.hx;
      float V = fmaf(0.5f * p.trap * (1.0f + p.e) * xr, xr, vcol) + p.k_int * (x[j].x * x[j].x + x[j].y * x[j].y);
      if (has_light) V = fmaf(gx[rl], gyv, V);
      const float a = V * dt;
      // exp(b dt_c) = exp(a ts_im) (cos(-a ts_re) + i sin(-a ts_re)); pure imaginary / pure real time
      // (the two cases the reference uses) need one or two MUFU operations instead of three
      if (pure_imag) {
        const float m = __expf(a * p.ts_im);
        x[j] = make_float2(x[j].x * m, x[j].y * m);
      } else {
        const float ph = -a * p.ts_re;
        float s, c;
        __sincosf(ph - 6.283185307179586f * rintf(ph * 0.15915494309189535f), &s, &c);
        if (!pure_real) {
          const float m = __expf(a * p.ts_im);
          s *= m;
          c *= m;
        }
        x[j] = cmul(x[j], make_float2(c, s));
      }
      acc = fmaf(x[j].x, x[j].x, fmaf(x[j].y, x[j].y, acc));
    }
    // CTA sum, then one float to each CTA of the cluster (own slot included)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid < kClCtas) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kClThreads / 32; ++w) t += red[w];
      st_cluster_f32(part_saddr + (uint32_t)(((k & 1) * kClCtas + (int)q) * sizeof(float)), (uint32_t)tid, t);
    }
    cluster_sync_all();  // partial sums of this step visible everywhere; also orders the reuse of `red`
    const float tot = (part[k & 1][0] + part[k & 1][1]) + (part[k & 1][2] + part[k & 1][3]);
    const float scale = rsqrtf(tot * p.dx * p.dx);  // solvers.py:111
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = make_float2(x[j].x * scale, x[j].y * scale);
  }
  float2* dst = reinterpret_cast<float2*>(p.y1) + (size_t)env * kClN * kClN + (size_t)row0 * kClN;
#pragma unroll
  for (int j = 0; j < 32; ++j) dst[j * kClThreads + tid] = x[j];
  cluster_sync_all();  // no CTA exits while a peer may still write into its shared memory
}

}  // namespace pdeopt
