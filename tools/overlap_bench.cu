// Can the shared-memory traffic of one CTA overlap the FP32 arithmetic of a co-resident CTA?
// Two 256-thread CTAs per SM: one runs packed butterflies only, the other bursts of 32 STS.64 +
// barrier + 32 LDS.64 only (the access pattern of an FFT exchange), then both do both.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "regfft.cuh"
using namespace pdeopt;
__global__ void __launch_bounds__(256, 2) k(float2* out, int iters, int mode) {
  extern __shared__ __align__(1024) unsigned char raw[];
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(raw) + threadIdx.x * 8;
  float2 x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = make_float2(1e-3f * (threadIdx.x + i), 1e-3f * i);
  // role: mode 0 = every CTA does arithmetic + exchange; 1 = arithmetic only; 2 = exchange only;
  // 3 = even CTAs arithmetic only, odd CTAs exchange only
  const bool do_fma = mode == 0 || mode == 1 || (mode == 3 && (blockIdx.x & 1) == 0);
  const bool do_lsu = mode == 0 || mode == 2 || (mode == 3 && (blockIdx.x & 1) == 1);
  for (int it = 0; it < iters; ++it) {
    if (do_fma) {
      DitF<32, 1, false>::run(x);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = mul2(x[i], make_float2(0.17f, 0.17f));
    }
    if (do_lsu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(base + i * 2048), "f"(x[i].x), "f"(x[i].y) : "memory");
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x[i].x), "=f"(x[i].y) : "r"((base ^ 8) + i * 2048));
      __syncthreads();
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 32; ++i) s = add2(s, x[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float2* out; cudaMalloc(&out, 296 * 256 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int iters = 2000;
  const char* names[] = {"both CTAs: arith + exchange", "both CTAs: arith only", "both CTAs: exchange only", "one CTA arith, one CTA exchange"};
  for (int mode = 0; mode < 4; ++mode) {
    k<<<296, 256, 65536>>>(out, iters, mode);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<296, 256, 65536>>>(out, iters, mode); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-34s %8.1f SM-cycles per iteration (%s)\n", names[mode], ms * 1e-3 * 1.965e9 / iters, cudaGetErrorString(cudaGetLastError()));
  }
}
