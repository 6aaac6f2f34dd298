// Strang split-step for complex fields that do not fit one SM (256x256 complex64 = 512 KB, BASELINE
// config 3) and, generally, any power-of-two grid up to 512x512: a multi-kernel path on the line-FFT
// engine with the state resident in L2 between kernels (128 envs x 512 KB = 64 MB of the 126 MB L2).
//
// Replaces StrangSplitting.step (pde_opt/numerics/solvers.py:99-122) with GPE2DTSControl.B_terms
// (pde_opt/numerics/equations/gross_pitaevskii.py:67-75); per step (A_term != 0):
//   K1 rows  fwd                               y0 -> W
//   K2 cols  fwd, * exp(A dt_c/2)/N^2, inv      W  -> W          (table in position order)
//   K3 rows  inv, * exp(b(psi0) dt_c), sum|.|^2 W  -> W, norm[env]   (b at psi0: solvers.py:109)
//   K4 rows  fwd of W / sqrt(norm dx^2)         W  -> W          (solvers.py:111 folded into the load)
//   K5 cols  fwd, * exp(A dt_c/2)/N^2, inv      W  -> W
//   K6 rows  inv                                W  -> y1
// With A_term == 0 (as shipped, gross_pitaevskii.py:62) the FFT round trips are identities and the
// step is two pointwise kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "linefft.cuh"

namespace pdeopt {

struct GpeLinesConst {
  int nx, ny, log2nx;
  float lo_x, lo_y, hx, hy;
  float trap, e, k_int;
  float ts_re, ts_im;
  const float* ctrl;  // [batch][8] or null: [1] amp [2] x0 [3] y0 [4] width of a Gaussian light spot
  // caller-evaluated additive potential `lights(t, x, y)` (gross_pitaevskii.py:61,72) for callables
  // outside the enumerated family: [nx][ny] floats per environment (stride 0 = shared), or null
  const float* light;
  long long light_env_stride;
};

// exp(b(psi0) dt_c) for one point: b = -i V, V = trap/2((1+e)x^2 + (1-e)y^2) + lights + k|psi0|^2
__device__ __forceinline__ float2 gpe_potential_factor(const GpeLinesConst& c, int env, int r, int col, float2 psi0, float dt) {
  const float xr = c.lo_x + (r + 0.5f) * c.hx, yc = c.lo_y + (col + 0.5f) * c.hy;
  float V = 0.5f * c.trap * ((1.0f + c.e) * xr * xr + (1.0f - c.e) * yc * yc) + c.k_int * (psi0.x * psi0.x + psi0.y * psi0.y);
  if (c.ctrl != nullptr) {
    const float* cc = c.ctrl + (size_t)env * 8;
    if (cc[1] != 0.f) {
      const float dx = xr - cc[2], dy = yc - cc[3];
      V += cc[1] * expf(-(dx * dx + dy * dy) * 0.5f / (cc[4] * cc[4]));
    }
  }
  if (c.light != nullptr) V += __ldg(c.light + (size_t)env * c.light_env_stride + (size_t)r * c.ny + col);
  const float a = V * dt;
  if (c.ts_re == 0.f) return make_float2(__expf(a * c.ts_im), 0.f);  // imaginary time: real factor
  const float ph = -a * c.ts_re;
  float s, cth;
  __sincosf(ph - 6.283185307179586f * rintf(ph * 0.15915494309189535f), &s, &cth);
  if (c.ts_im == 0.f) return make_float2(cth, s);                     // real time: pure phase
  const float m = __expf(a * c.ts_im);
  return make_float2(m * cth, m * s);
}

__device__ __forceinline__ float lf_block_sum(float v) {
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// complex multiplier table laid out like one env's field: tab[pos_row * ny + pos_col]
struct LfMidCTab {
  const float2* tab;
  LineGeom g;  // the column-pass geometry with outer = 0: the table is shared by the batch
  __device__ __forceinline__ LineGeom gaux() const { return g; }
  __device__ __forceinline__ float2 apply(float2 v, long long off_aux, long long, int) const { return cmul(v, tab[off_aux]); }
  using Aux = float2;
  __device__ __forceinline__ float2 fetch(long long off_aux, long long) const { return __ldg(tab + off_aux); }
  __device__ __forceinline__ float2 apply_aux(float2 v, float2 t, long long, int) const { return cmul(v, t); }
};

// K3 storer: rows inverse done -> multiply by the potential factor (b at psi0), accumulate the norm.
struct LfStorePotential {
  float2* out;
  const float2* psi0;
  float* norm;  // [batch]
  GpeLinesConst c;
  float dt;
  float acc;
  LineGeom g;
  __device__ __forceinline__ LineGeom gout() const { return g; }
  __device__ __forceinline__ float2 pre(long long o) const { return psi0[o]; }
  __device__ __forceinline__ void store(long long o, long long line, int idx, float2 v, float2 p0) {
    const int l32 = (int)line, env = l32 >> c.log2nx, r = l32 & (c.nx - 1);
    const float2 w = cmul(v, gpe_potential_factor(c, env, r, idx, p0, dt));
    out[o] = w;
    acc = fmaf(w.x, w.x, fmaf(w.y, w.y, acc));
  }
  __device__ __forceinline__ void flush(long long l0) {
    const float t = lf_block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(norm + (l0 >> c.log2nx), t);
    acc = 0.f;
  }
};

// Fused K3+K4 (rows): inverse, times exp(b(psi0) dt_c) with the norm accumulated, forward again.  The
// L2 renormalisation (solvers.py:111) is a scalar per environment and commutes with the transforms,
// so it is applied by the following column pass (LfMidCTabScaled) instead of a pass of its own.
struct LfMidPotentialIMF {
  const float2* psi0;
  float* norm;  // [batch]
  GpeLinesConst c;
  float dt;
  float acc;
  __device__ __forceinline__ LineGeom gaux() const { return LineGeom{1, 1, 0, 0, 1, 0, 0}; }
  __device__ __forceinline__ float2 pre(long long o) const { return psi0[o]; }
  __device__ __forceinline__ float2 apply(float2 v, long long, long long line, int idx, float2 p0) {
    const int l32 = (int)line, env = l32 >> c.log2nx, r = l32 & (c.nx - 1);
    const float2 w = cmul(v, gpe_potential_factor(c, env, r, idx, p0, dt));
    acc = fmaf(w.x, w.x, fmaf(w.y, w.y, acc));
    return w;
  }
  __device__ __forceinline__ void flush(long long l0) {
    const float t = lf_block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(norm + (l0 >> c.log2nx), t);
    acc = 0.f;
  }
};
// Fused K6+K1 (rows): inverse -> y1 (the state the next step's b(psi0) needs, and the output) -> forward
struct LfMidStoreState {
  float2* y1;
  __device__ __forceinline__ LineGeom gaux() const { return LineGeom{1, 1, 0, 0, 1, 0, 0}; }
  __device__ __forceinline__ float2 pre(long long) const { return make_float2(0.f, 0.f); }
  __device__ __forceinline__ float2 apply(float2 v, long long o, long long, int, float2) const {
    y1[o] = v;
    return v;
  }
  __device__ __forceinline__ void flush(long long) {}
};
// column pass multiplier exp(A dt_c / 2) / N^2 times the pending renormalisation 1/sqrt(norm dx^2)
struct LfMidCTabScaled {
  const float2* tab;
  const float* norm;
  LineGeom g;  // column-pass geometry with outer = 0 (table shared by the batch)
  int log2ny;
  float dx2;
  __device__ __forceinline__ LineGeom gaux() const { return g; }
  __device__ __forceinline__ float2 apply(float2 v, long long off_aux, long long line, int) const {
    const float s = rsqrtf(norm[(int)line >> log2ny] * dx2);
    const float2 t = tab[off_aux];
    return cmul(v, make_float2(t.x * s, t.y * s));
  }
  using Aux = float2;
  __device__ __forceinline__ float2 fetch(long long off_aux, long long) const { return __ldg(tab + off_aux); }
  __device__ __forceinline__ float2 apply_aux(float2 v, float2 t, long long line, int) const {
    const float s = rsqrtf(norm[(int)line >> log2ny] * dx2);
    return cmul(v, make_float2(t.x * s, t.y * s));
  }
};

// K4 loader: W / sqrt(norm dx^2)
struct LfLoadNormalised {
  const float2* p;
  const float* norm;
  int nx, ny, log2nx;
  float dx2;
  LineGeom g;
  __device__ __forceinline__ LineGeom gin() const { return g; }
  __device__ __forceinline__ float2 load(long long off, long long line, int) const {
    const float s = rsqrtf(norm[(int)line >> log2nx] * dx2);
    const float2 v = p[off];
    return make_float2(v.x * s, v.y * s);
  }
};

// exp(A_term * 0.5 * dt_c) / (nx ny) in position order
__global__ void strang_lines_etab_kernel(const float2* __restrict__ a_term, float2* __restrict__ etab, int nx, int ny,
                                         float hr, float hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx * ny) return;
  const int pr = i / ny, pc = i % ny;
  const float2 a = a_term[(size_t)line_pos_to_freq(nx, pr) * ny + line_pos_to_freq(ny, pc)];
  const float re = a.x * hr - a.y * hi, im = a.x * hi + a.y * hr;
  const float m = expf(re) / float(nx * ny);
  float s, c;
  sincosf(im, &s, &c);
  etab[i] = make_float2(m * c, m * s);
}

// A_term == 0 (the equation as shipped, gross_pitaevskii.py:62): the step is pointwise except for the
// renormalisation, y = psi0 * exp(b(psi0) dt_c) / ||.||.  One kernel per step: the division by the
// norm of step k-1 is deferred to the load of step k (psi0 = w_{k-1} / ||w_{k-1}||, `prev_norm`), so
// each step reads and writes the state once; a final scale kernel applies the last norm.
__global__ void __launch_bounds__(256) strang_lines_potential_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                                     const float* __restrict__ prev_norm,
                                                                     float* __restrict__ norm, GpeLinesConst c, float dt,
                                                                     float dx2, int blocks_per_env) {
  const int env = blockIdx.x / blocks_per_env, blk = blockIdx.x % blocks_per_env;
  const int npts = c.nx * c.ny;
  const int per = (npts + blocks_per_env - 1) / blocks_per_env;
  const int beg = blk * per, end = min(npts, beg + per);
  const float s = prev_norm ? rsqrtf(prev_norm[env] * dx2) : 1.0f;
  float acc = 0.f;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const size_t o = (size_t)env * npts + i;
    float2 p0 = in[o];
    p0 = make_float2(p0.x * s, p0.y * s);
    const float2 w = cmul(p0, gpe_potential_factor(c, env, i / c.ny, i % c.ny, p0, dt));
    out[o] = w;
    acc = fmaf(w.x, w.x, fmaf(w.y, w.y, acc));
  }
  const float t = lf_block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(norm + env, t);
}
__global__ void __launch_bounds__(256) strang_lines_scale_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                                 const float* __restrict__ norm, int npts, float dx2,
                                                                 long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float s = rsqrtf(norm[i / npts] * dx2);
  const float2 v = in[i];
  out[i] = make_float2(v.x * s, v.y * s);
}

}  // namespace pdeopt
