// How much shared-memory bandwidth can ONE CTA drive during an FFT-style exchange (store burst,
// barrier, load burst, barrier)?  Varies threads per CTA and access width.  One CTA per SM.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
template <int W>  // W = 8 or 16 bytes per access; each thread moves 256 bytes per burst
__global__ void k(float* out, int iters) {
  extern __shared__ __align__(1024) unsigned char raw[];
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(raw) + threadIdx.x * W;
  float x[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) x[i] = threadIdx.x + i;
  constexpr int N = 256 / W;
  const uint32_t stride = blockDim.x * W;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (W == 8) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(base + i * stride), "f"(x[2 * i]), "f"(x[2 * i + 1]) : "memory");
      else asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(base + i * stride), "f"(x[4 * i]), "f"(x[4 * i + 1]), "f"(x[4 * i + 2]), "f"(x[4 * i + 3]) : "memory");
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (W == 8) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "r"((base ^ (W * 2)) + i * stride));
      else asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[4 * i]), "=f"(x[4 * i + 1]), "=f"(x[4 * i + 2]), "=f"(x[4 * i + 3]) : "r"((base ^ (W * 2)) + i * stride));
    }
    __syncthreads();
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int W> void run(int threads, int ctas_per_sm, float* out) {
  const int iters = 2000, smem = threads * 256;
  cudaFuncSetAttribute(k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<W><<<148 * ctas_per_sm, threads, smem>>>(out, iters);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<W><<<148 * ctas_per_sm, threads, smem>>>(out, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9 / iters;
  const double bytes = 2.0 * threads * 256 * ctas_per_sm;
  printf("%3d threads x %d CTA/SM, %2d-byte accesses: %7.1f cycles per exchange, %6.1f B/clk/SM (%s)\n", threads, ctas_per_sm, W, cyc, bytes / cyc,
         cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 1024 * 4);
  run<8>(256, 1, out); run<16>(256, 1, out); run<8>(512, 1, out); run<16>(512, 1, out); run<8>(256, 2, out); run<16>(256, 2, out); run<8>(128, 1, out); run<8>(128, 4, out);
}
