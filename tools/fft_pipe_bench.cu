// How close can register butterflies get to the FP32 pipe peak?  Pure-register 32-point DIT
// transforms in a loop (no shared memory, no barriers), 4 warps per scheduler.
//   nvcc -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -O3 -I pde_opt_b200/csrc -o tools/fft_pipe_bench tools/fft_pipe_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "regfft.cuh"
using namespace pdeopt;

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float2* out, long long* cyc, int iters) {
  float2 x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = make_float2(1e-3f * (threadIdx.x + i), 1e-3f * i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) DitF<32, 1, false>::run(x);
    if (MODE == 1) Dif<32, 1, false>::run(x);
    if (MODE == 2) { DitF<16, 1, false>::run(x); DitF<16, 1, false>::run(x + 16); }
    if (MODE == 3) { static_for<0, 4>([&](auto s) { DitF<8, 1, false>::run(x + 8 * decltype(s)::value); }); }
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = mul2(x[i], make_float2(0.17f, 0.17f));
  }
  long long t1 = clock64();
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 32; ++i) s = add2(s, x[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, float2* out, long long* cyc) {
  const int iters = 2000;
  for (int grid : {148, 296}) {
    k<MODE><<<grid, 256>>>(out, cyc, iters);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, cyc, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-10s grid %d: CTA0 %lld cycles (%.1f / iteration); kernel %.3f ms = %.1f SM-cycles per iteration at 1.965 GHz\n", name, grid, c,
           (double)c / iters, ms, ms * 1e-3 * 1.965e9 / iters);
  }
}
int main() {
  float2* out; long long* cyc;
  cudaMalloc(&out, 296 * 256 * 8); cudaMalloc(&cyc, 296 * 8);
  run<0>("DitF32", out, cyc); run<1>("Dif32", out, cyc); run<2>("2xDitF16", out, cyc); run<3>("4xDitF8", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
