// Shared-memory instruction throughput on B200 (sm_100a): LDS/STS .64 and .128, contiguous and
// half-warp-swizzled address patterns, with inline PTX so the access width is what is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/smem_bench tools/smem_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int ITERS = 2048;
enum Op { LD64_CONTIG, LD64_HALFWARP, LD64_STRIDE16, LD128_CONTIG, LD128_XOR, ST64_CONTIG, ST128_CONTIG, LD64_ST64 };

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  uint32_t a;
  if (OP == LD64_CONTIG || OP == ST64_CONTIG || OP == LD64_ST64) a = base + warp * 2048 + lane * 8;
  else if (OP == LD64_HALFWARP) a = base + warp * 2048 + ((lane & 15) * 8) + (lane >> 4) * 1024 + ((lane & 15) ^ 5) * 0;  // two half-warps in different 128 B lines
  else if (OP == LD64_STRIDE16) a = base + warp * 2048 + (lane & 15) * 16 + (lane >> 4) * 8;  // every lane distinct 8 B slot, 2 lines
  else if (OP == LD128_CONTIG || OP == ST128_CONTIG) a = base + warp * 2048 + lane * 16;
  else a = base + warp * 2048 + ((lane * 16) ^ ((warp & 7) << 4));
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t ai = a + ((i & 3) * 512);
      if (OP == LD64_CONTIG || OP == LD64_HALFWARP || OP == LD64_STRIDE16) {
        float x, y;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(ai));
        acc0 += x; acc1 += y;
      } else if (OP == LD128_CONTIG || OP == LD128_XOR) {
        float x, y, z, w;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(ai));
        acc0 += x; acc1 += y; acc2 += z; acc3 += w;
      } else if (OP == ST64_CONTIG) {
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(ai), "f"(acc0), "f"(acc1) : "memory");
        acc0 += 1.f;
      } else if (OP == ST128_CONTIG) {
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(ai), "f"(acc0), "f"(acc1), "f"(acc2), "f"(acc3) : "memory");
        acc0 += 1.f;
      } else {
        float x, y;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(ai));
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(ai ^ 1024), "f"(acc0), "f"(acc1) : "memory");
        acc0 += x; acc1 += y;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + tid] = acc0 + acc1 + acc2 + acc3;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int bytes_per_thread_inst, float* out, long long* cyc) {
  cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<OP><<<148, 512, 65536>>>(out, cyc);
  k<OP><<<148, 512, 65536>>>(out, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double warp_inst = (double)ITERS * 8 * 16 * (OP == LD64_ST64 ? 2 : 1);
  printf("%-16s cycles=%9lld  %6.3f warp-inst/clk/SM  %7.1f B/clk/SM\n", name, c, warp_inst / c, warp_inst * 32 * bytes_per_thread_inst / c);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  run<LD64_CONTIG>("LDS.64 contig", 8, out, cyc);
  run<LD64_HALFWARP>("LDS.64 halfwarp", 8, out, cyc);
  run<LD64_STRIDE16>("LDS.64 stride16", 8, out, cyc);
  run<LD128_CONTIG>("LDS.128 contig", 16, out, cyc);
  run<LD128_XOR>("LDS.128 xor", 16, out, cyc);
  run<ST64_CONTIG>("STS.64 contig", 8, out, cyc);
  run<ST128_CONTIG>("STS.128 contig", 16, out, cyc);
  run<LD64_ST64>("LDS.64+STS.64", 8, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
