// DSMEM read bandwidth inside a cluster of 4 CTAs (B200): every thread reads 32 float2 "slots" from
// each of the 3 peer CTAs (the access pattern of a radix-2x2 exchange of a 128 KB field), repeated.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_bench tools/dsmem_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 512, kSlots = 32;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1) dsmem_kernel(float2* out, int reps, long long* cycles) {
  extern __shared__ __align__(16) float2 W[];
  const int tid = threadIdx.x;
  const uint32_t q = ctarank();
  for (int n = 0; n < kSlots; ++n) W[n * kThreads + tid] = make_float2(float(q), float(n));
  csync();
  float2 acc = make_float2(0.f, 0.f);
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(W);
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int peer = 1; peer < 4; ++peer) {
      uint32_t raddr;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(base), "r"((q + peer) & 3));
#pragma unroll
      for (int n = 0; n < kSlots; ++n) {
        float2 v;
        asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(raddr + (uint32_t)((n * kThreads + tid) * 8)));
        acc.x += v.x;
        acc.y += v.y;
      }
    }
  }
  const long long t1 = clock64();
  csync();
  out[blockIdx.x * kThreads + tid] = acc;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  const int clusters = 33, grid = clusters * 4, reps = 64;
  float2* out; long long* cyc;
  cudaMalloc(&out, sizeof(float2) * grid * kThreads);
  cudaMalloc(&cyc, sizeof(long long) * grid);
  const size_t smem = sizeof(float2) * kSlots * kThreads;
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int it = 0; it < 2; ++it) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_kernel, out, reps, cyc);
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return 1; }
  }
  long long h[grid];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  const double bytes = 3.0 * kSlots * kThreads * 8 * reps;  // per CTA
  printf("DSMEM remote reads: %.0f cycles per CTA for %.1f MB -> %.1f B/clk/SM (4-CTA clusters, %d clusters)\n", avg, bytes / 1e6, bytes / avg, clusters);
  return 0;
}
