// Pipe-throughput microbenchmarks for B200 (sm_100a): FP32 scalar vs packed f32x2, MUFU, SHFL,
// shared-memory bandwidth, and TMEM park ld/st.  Gives the measured FFMA peak that the fused
// stepper's roofline fraction is quoted against (BASELINE.md section 2 asks for it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

enum Op { FFMA, FADD, FMUL, FFMA2, FADD2, MIX_FMA_ADD, MUFU_EX2, MUFU_LG2, MUFU_RCP, SHFL, LDS64, LDS128, STS64, TMEM_LD, TMEM_ST };

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cycles, int nwarps_active) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  float a[8];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 1.0f + tid * 1e-6f + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
  const float b = 0.999f, c = 1e-3f;
  const unsigned long long pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
  const unsigned long long pc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
  for (int i = tid; i < 8192; i += blockDim.x) smem[i] = i;
  __shared__ uint32_t tbase;
  if (OP == TMEM_LD || OP == TMEM_ST) {
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int warp = tid >> 5;
  uint32_t taddr = 0;
  if (OP == TMEM_LD || OP == TMEM_ST) taddr = tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) & 7) * 64u;
  long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int it = 0; it < ITERS; ++it) {
      if (OP == FFMA) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
      } else if (OP == FADD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a[i] + c;
      } else if (OP == FMUL) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a[i] * b;
      } else if (OP == FFMA2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
      } else if (OP == FADD2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
      } else if (OP == MIX_FMA_ADD) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) { a[i] = fmaf(a[i], b, c); a[i + 1] = a[i + 1] + c; }
#pragma unroll
        for (int i = 0; i < 8; i += 2) { a[i] = fmaf(a[i], b, c); a[i + 1] = a[i + 1] + c; }
      } else if (OP == MUFU_EX2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = exp2f(a[i]) * 1e-3f;  // ex2.approx under -use_fast_math-less: use intrinsic below
      } else if (OP == MUFU_LG2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("lg2.approx.f32 %0, %0;" : "+f"(a[i]));
      } else if (OP == MUFU_RCP) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("rcp.approx.f32 %0, %0;" : "+f"(a[i]));
      } else if (OP == SHFL) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __shfl_sync(0xffffffffu, a[i], (tid + 1) & 31);
      } else if (OP == LDS64) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(smem + ((tid * 2 + i * 2048 + it * 64) & 8191 & ~1)); a[i] += v.x + v.y; }
      } else if (OP == LDS128) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float4 v = *reinterpret_cast<float4*>(smem + ((tid * 4 + i * 1024 + it * 128) & 8191 & ~3)); a[i] += v.x + v.w; }
      } else if (OP == STS64) {
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float2*>(smem + ((tid * 2 + i * 2048 + it * 64) & 8191 & ~1)) = make_float2(a[i], a[i]);
      } else if (OP == TMEM_ST) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%2,%2,%3,%3,%4,%4,%5,%5,%6,%6,%7,%7,%8,%8};" ::"r"(taddr + i * 16),
                       "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else if (OP == TMEM_LD) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float r[16];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
                         "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]) : "r"(taddr + i * 16) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          a[i] += r[0] + r[15];
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  __syncthreads();
  if ((OP == TMEM_LD || OP == TMEM_ST) && tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

// TMEM park correctness: every thread stores 64 distinct values, reads them back.
__global__ void __launch_bounds__(512, 1) tmem_park_check(int* bad) {
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;
  for (int ch = 0; ch < 4; ++ch) {
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = float(tid * 64 + ch * 16 + i + blockIdx.x * 100000);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr + ch * 16),
                 "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
                 "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  int nbad = 0;
  for (int ch = 0; ch < 4; ++ch) {
    float r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
                   "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]) : "r"(taddr + ch * 16) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) nbad += (r[i] != float(tid * 64 + ch * 16 + i + blockIdx.x * 100000));
  }
  if (nbad) atomicAdd(bad, nbad);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase));
}

template <int OP>
int run(const char* name, double ops_per_inst_per_thread, int threads, const char* unit) {
  int nsm = 148;
  float* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(float) * nsm * 1024));
  CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
  CK(cudaFuncSetAttribute(bench<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int rep = 0; rep < 2; ++rep) bench<OP><<<nsm, threads, 200 * 1024>>>(out, cyc, threads / 32);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(nsm);
  CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
  double avg = 0; for (auto c : h) avg += c; avg /= nsm;
  int per_iter = (OP == TMEM_LD || OP == TMEM_ST) ? 4 : 8;
  double inst = (double)ITERS * per_iter * threads;  // thread-instructions per SM
  printf("%-14s threads=%4d  cycles=%9.0f  %8.2f thread-inst/clk/SM  -> %8.2f %s/clk/SM\n", name, threads, avg, inst / avg, inst / avg * ops_per_inst_per_thread, unit);
  cudaFree(out); cudaFree(cyc);
  return 0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s  SMs=%d  clock=%d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  for (int threads : {512, 1024}) {
    run<FFMA>("FFMA", 2, threads, "flop");
    run<FADD>("FADD", 1, threads, "flop");
    run<FMUL>("FMUL", 1, threads, "flop");
    run<FFMA2>("FFMA2", 4, threads, "flop");
    run<FADD2>("FADD2", 2, threads, "flop");
    run<MIX_FMA_ADD>("FFMA+FADD", 1.5, threads, "flop");
    run<MUFU_LG2>("MUFU.LG2", 1, threads, "op");
    run<MUFU_RCP>("MUFU.RCP", 1, threads, "op");
    run<SHFL>("SHFL", 4, threads, "B");
    run<LDS64>("LDS.64", 8, threads, "B");
    run<LDS128>("LDS.128", 16, threads, "B");
    run<STS64>("STS.64", 8, threads, "B");
    run<TMEM_ST>("TMEM.ST.x16", 64, threads, "B");
    run<TMEM_LD>("TMEM.LD.x16", 64, threads, "B");
  }
  int* bad; CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
  tmem_park_check<<<296, 512>>>(bad);
  CK(cudaDeviceSynchronize());
  int hb = -1; CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
  printf("tmem_park_check mismatches: %d\n", hb);
  return 0;
}
