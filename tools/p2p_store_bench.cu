// How fast can SM stores push data into a PEER GPU's memory over NVLink, as a function of the contiguous run length
// per half-warp and of the bytes per thread?  (B200 x2, sm_100a.)  Motivation: the slab transposes fused into the line-FFT
// kernels (pdeopt_fft_lines_to_peers) reach 270-390 GB/s per rank, well under the ~770 GB/s a peer cudaMemcpy gets.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2p_store_bench tools/p2p_store_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// every group of (run_bytes / VEC) consecutive threads writes one contiguous run; consecutive runs are `stride_bytes` apart
// (wrapping inside the buffer), so run_bytes == stride_bytes is a plain streaming write
template <int VEC>
__global__ void store_kernel(char* __restrict__ dst, size_t total_bytes, int run_bytes, size_t stride_bytes, size_t buf_bytes) {
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  const int per_run = run_bytes / VEC;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i * VEC < total_bytes; i += nthreads) {
    const size_t run = i / per_run, within = i % per_run;
    const size_t off = (run * stride_bytes) % buf_bytes + within * VEC;
    if (VEC == 8) *reinterpret_cast<float2*>(dst + off) = make_float2(1.f, 2.f);
    else *reinterpret_cast<float4*>(dst + off) = make_float4(1.f, 2.f, 3.f, 4.f);
  }
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t buf = (size_t)512 << 20, total = (size_t)256 << 20;
  char *local, *peer;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, buf));
  CK(cudaSetDevice(0)); CK(cudaMalloc(&local, buf));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // reference: copy engine
  CK(cudaMemcpyPeer(peer, 1, local, 0, total));
  CK(cudaEventRecord(e0)); for (int r = 0; r < 5; ++r) CK(cudaMemcpyPeerAsync(peer, 1, local, 0, total, 0)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("cudaMemcpyPeer            : %7.1f GB/s\n", 5.0 * total / ms / 1e6);
  const int runs[] = {64, 128, 256, 512, 2048};
  for (int tgt = 0; tgt < 2; ++tgt) {
    char* dst = tgt ? peer : local;
    for (int vec : {8, 16}) for (int rb : runs) for (int scattered = 0; scattered < 2; ++scattered) for (int ctas : {148 * 2, 148 * 8}) {
      if (rb < vec * 4) continue;
      const size_t stride = scattered ? (size_t)rb * 257 : (size_t)rb;  // scattered: runs 257 run-lengths apart
      auto launch = [&]() {
        if (vec == 8) store_kernel<8><<<ctas, 256>>>(dst, total, rb, stride, buf);
        else store_kernel<16><<<ctas, 256>>>(dst, total, rb, stride, buf);
      };
      launch(); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); for (int r = 0; r < 3; ++r) launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("%s  %2d B/thread  run %4d B  %s  %4d CTAs : %7.1f GB/s\n", tgt ? "peer " : "local", vec, rb, scattered ? "scattered " : "contiguous", ctas,
             3.0 * total / ms / 1e6);
    }
  }
  return 0;
}
