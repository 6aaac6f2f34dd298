// Isolated timing of the three register passes of rfft128.cuh (two 256-thread CTAs per SM, as in
// the stepper): which pass runs far from the FP32-pipe bound, and why.
//   nvcc -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -O3 -I pde_opt_b200/csrc -o tools/pass_bench tools/pass_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "rfft128.cuh"
using namespace pdeopt;
using namespace pdeopt::rf;

struct __align__(1024) Smem {
  float2 W[kRows * kH];
  float4 T[kTRows * kTCols];
  float2 twb[8 * 16];
  float2 tw64[32];
};

template <int MODE>
__global__ void __launch_bounds__(kThreadsR, 2) k(float2* out, long long* cyc, int iters) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& S = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < kRows * kH; i += kThreadsR) S.W[i] = make_float2(1e-3f * (i & 255), 1e-3f * (i & 127));
  for (int i = tid; i < kTRows * kTCols; i += kThreadsR) S.T[i] = make_float4(0.6f / 8192, 0.6f / 8192, 0.3f / 8192, 0.3f / 8192);
  if (tid < 128) { float s, c; sincospif(-2.0f * float((tid & 7) * (tid >> 3)) / 128.0f, &s, &c); S.twb[tid] = make_float2(c, s); }
  if (tid < 32) { float s, c; sincospif(-2.0f * float(tid) / 64.0f, &s, &c); S.tw64[tid] = make_float2(c, s); }
  __syncthreads();
  const RFft F((uint32_t)__cvta_generic_to_shared(S.W), (uint32_t)__cvta_generic_to_shared(S.T), tid);
  float2 x[32];
  gather_nat(F, x);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { passA_fwd(F, x); __syncthreads(); passA_inv(F, x); __syncthreads(); }
    if (MODE == 1) { passB_fwd(F, S.twb, S.tw64, x); __syncthreads(); passB_inv(F, S.twb, S.tw64, x); __syncthreads(); }
    if (MODE == 2) { passC_filter(F, x); __syncthreads(); }
    if (MODE == 3) { scatter_nat(F, x); __syncthreads(); gather_nat<true>(F, x); __syncthreads(); }
    if (MODE == 4) {  // the whole transform chain of one step
      passA_fwd(F, x); __syncthreads(); passB_fwd(F, S.twb, S.tw64, x); __syncthreads(); passC_filter(F, x); __syncthreads();
      passB_inv(F, S.twb, S.tw64, x); __syncthreads(); passA_inv(F, x); __syncthreads();
    }
    if (MODE != 3) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = mul2(x[i], make_float2(0.01f, 0.01f));
    }
  }
  long long t1 = clock64();
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 32; ++i) s = add2(s, x[i]);
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, float2* out, long long* cyc) {
  const int iters = 400;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  for (int grid : {148, 296}) {
    k<MODE><<<grid, kThreadsR, sizeof(Smem)>>>(out, cyc, iters);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, kThreadsR, sizeof(Smem)>>>(out, cyc, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-14s grid %3d: %8.1f SM-cycles per iteration per CTA-slot (kernel %.3f ms) %s\n", name, grid,
           ms * 1e-3 * 1.965e9 / iters / (grid / 148), ms, cudaGetErrorString(cudaGetLastError()));
  }
}
int main() {
  float2* out; long long* cyc;
  cudaMalloc(&out, 296 * 256 * 8); cudaMalloc(&cyc, 296 * 8);
  run<0>("A fwd+inv", out, cyc); run<1>("B fwd+inv", out, cyc); run<2>("C+filter", out, cyc); run<3>("scatter+gather", out, cyc); run<4>("all passes", out, cyc);
}
