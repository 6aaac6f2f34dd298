// Quantised-vortex detection for batched GPE states: phase circulation on every grid cell.
//
// Replaces pde_opt/rl_utils.py:19-84 (detect_vortices), the reward helper of the GPE environments,
// as a fused epilogue: one pass over psi gives the integer winding of every plaquette and the
// per-environment counts, so a reward like "number of vortices" needs no host round trip of the state.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdeopt {

struct VortexParams {
  const float2* psi;  // [batch][n0][n1] complex
  int32_t* winding;   // [batch][n0][n1] or null
  int32_t* counts;    // [batch][3]: number of cells with non-zero winding, sum, sum of |winding| (zeroed by the caller)
  int n0, n1, batch;
  float amp_thresh, tol;
};

// rl_utils.py:15-17: map to [-pi, pi) with the sign convention of Python's %
__device__ __forceinline__ float vx_wrap(float x) {
  const float two_pi = 6.28318530717958647692f, pi = 3.14159265358979323846f;
  float y = x + pi;
  y -= two_pi * floorf(y / two_pi);
  return y - pi;
}

// One thread per cell of a 32 x 8 tile; the phases of the (tile + 1)^2 corners are staged in shared memory.
__global__ void __launch_bounds__(256) vortex_kernel(const __grid_constant__ VortexParams p) {
  __shared__ float th[9][33];
  __shared__ float rho[9][33];
  __shared__ int red[3];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * 8, j0 = blockIdx.x * 32;
  const int tj = threadIdx.x & 31, ti = threadIdx.x >> 5;
  const float2* ps = p.psi + (size_t)b * p.n0 * p.n1;
  if (threadIdx.x < 3) red[threadIdx.x] = 0;
  for (int e = threadIdx.x; e < 9 * 33; e += 256) {
    const int a = e / 33, c = e - a * 33;
    int gi = i0 + a, gj = j0 + c;
    gi = gi >= p.n0 ? gi - p.n0 : gi;
    gj = gj >= p.n1 ? gj - p.n1 : gj;
    float t = 0.f, r = 0.f;
    if (i0 + a <= p.n0 && j0 + c <= p.n1) {
      const float2 v = ps[(size_t)gi * p.n1 + gj];
      t = atan2f(v.y, v.x);  // jnp.angle
      r = v.x * v.x + v.y * v.y;
    }
    th[a][c] = t;
    rho[a][c] = r;
  }
  __syncthreads();
  const int i = i0 + ti, j = j0 + tj;
  int n = 0;
  if (i < p.n0 && j < p.n1) {
    // rl_utils.py:46-56: edges of the plaquette with corners (i, j), (i, j+1), (i+1, j+1), (i+1, j)
    const float t00 = th[ti][tj], t01 = th[ti][tj + 1], t10 = th[ti + 1][tj], t11 = th[ti + 1][tj + 1];
    const float dx0 = vx_wrap(t01 - t00);  // dth_x[i][j]
    const float dy1 = vx_wrap(t11 - t01);  // dth_y[i][j+1]
    const float dx1 = vx_wrap(t11 - t10);  // dth_x[i+1][j]
    const float dy0 = vx_wrap(t10 - t00);  // dth_y[i][j]
    const float circ = dx0 + dy1 - dx1 - dy0;
    const float nf = circ / 6.28318530717958647692f;
    n = (fabsf(nf) >= p.tol) ? (int)rintf(nf) : 0;
    if (p.amp_thresh > 0.f) {
      const float rc = 0.25f * (rho[ti][tj] + rho[ti + 1][tj] + rho[ti][tj + 1] + rho[ti + 1][tj + 1]);
      if (!(rc >= p.amp_thresh)) n = 0;
    }
    if (p.winding) p.winding[((size_t)b * p.n0 + i) * p.n1 + j] = n;
  }
  // per-tile reduction, then three atomics per tile
  const unsigned nz = __ballot_sync(0xffffffffu, n != 0);
  int s = n, a = n < 0 ? -n : n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    a += __shfl_xor_sync(0xffffffffu, a, o);
  }
  if (tj == 0 && nz != 0) {
    atomicAdd(&red[0], __popc(nz));
    atomicAdd(&red[1], s);
    atomicAdd(&red[2], a);
  }
  __syncthreads();
  if (threadIdx.x < 3 && red[threadIdx.x] != 0) atomicAdd(p.counts + 3 * b + threadIdx.x, red[threadIdx.x]);
}

}  // namespace pdeopt
