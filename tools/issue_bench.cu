// Does a non-FMA instruction issue in the shadow of a packed f32x2 instruction (2 pipe cycles), or
// does the packed instruction hold the SMSP's dispatch port for both cycles?  B200 (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_bench tools/issue_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
constexpr int ITERS = 4096;
enum { P8, P8_LOP8, P8_IADD8, S8, S8_LOP8, P8_LDS4, P8_MUFU2, S16, P8_S8, P8_SHFL4, P8_STS4, F8, F8_LOP8, F8_LDS8, F8_MOV8, F8_MUFU4, F8_STS2 };
template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc) {
  __shared__ __align__(16) float sm[4096];
  const int tid = threadIdx.x;
  for (int i = tid; i < 4096; i += 512) sm[i] = i;
  __syncthreads();
  unsigned long long p[8];
  float a[16];
  uint32_t q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = (((unsigned long long)__float_as_uint(1.0f + i)) << 32) | __float_as_uint(0.5f + tid); q[i] = tid * 7 + i; }
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0f + tid * 1e-6f + i;
  const unsigned long long pc = (((unsigned long long)__float_as_uint(1e-3f)) << 32) | __float_as_uint(1e-3f);
  float b = 0.999f + tid * 1e-9f, c = 1e-3f + tid * 1e-9f;
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm) + tid * 8;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == P8 || OP == P8_LOP8 || OP == P8_IADD8 || OP == P8_LDS4 || OP == P8_MUFU2 || OP == P8_S8 || OP == P8_SHFL4 || OP == P8_STS4)
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
      if (OP == F8 || OP == F8_LOP8 || OP == F8_LDS8 || OP == F8_MOV8 || OP == F8_MUFU4 || OP == F8_STS2)
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(p[(i + 3) & 7]), "l"(p[(i + 5) & 7]));
      if (OP == F8_LOP8) asm volatile("xor.b32 %0, %0, %1;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]));
      if (OP == F8_MOV8) asm volatile("mov.b32 %0, %1;" : "=r"(q[i]) : "r"(q[(i + 1) & 7]));
      if (OP == F8_LDS8) { float x, y; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(sa + i * 512)); a[i] = x; a[8 + i] = y; }
      if (OP == F8_MUFU4 && (i & 1) == 0) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == F8_STS2 && (i & 3) == 0) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(sa + i * 512), "f"(a[i]), "f"(a[i + 1]) : "memory");
      if (OP == S8 || OP == S8_LOP8 || OP == S16 || OP == P8_S8) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
      if (OP == S16) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[8 + i]) : "f"(b), "f"(c));
      if (OP == P8_LOP8 || OP == S8_LOP8) asm volatile("xor.b32 %0, %0, %1;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]));
      if (OP == P8_IADD8) asm volatile("add.s32 %0, %0, %1;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]));
      if (OP == P8_LDS4 && (i & 1) == 0) { float x, y; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(sa + i * 512)); a[i] = x; a[i + 1] = y; }
      if (OP == P8_STS4 && (i & 1) == 0) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(sa + i * 512), "f"(a[i]), "f"(a[i + 1]) : "memory");
      if (OP == P8_MUFU2 && (i & 3) == 0) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == P8_SHFL4 && (i & 1) == 0) a[i] = __shfl_sync(0xffffffffu, a[i], (tid + 1) & 31);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32)) + a[i] + a[8 + i] + (float)q[i];
  out[blockIdx.x * 512 + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* n, float* out, long long* cyc) {
  k<OP><<<148, 512>>>(out, cyc); k<OP><<<148, 512>>>(out, cyc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %7.2f cycles per loop iteration per SMSP (4 warps)\n", n, (double)c / ITERS);
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  run<P8>("8 FADD2", out, cyc); run<P8_LOP8>("8 FADD2 + 8 LOP3", out, cyc); run<P8_IADD8>("8 FADD2 + 8 IADD", out, cyc);
  run<S8>("8 FFMA", out, cyc); run<S16>("16 FFMA", out, cyc); run<S8_LOP8>("8 FFMA + 8 LOP3", out, cyc); run<P8_S8>("8 FADD2 + 8 FFMA", out, cyc);
  run<P8_LDS4>("8 FADD2 + 4 LDS.64", out, cyc); run<P8_STS4>("8 FADD2 + 4 STS.64", out, cyc); run<P8_MUFU2>("8 FADD2 + 2 MUFU", out, cyc); run<P8_SHFL4>("8 FADD2 + 4 SHFL", out, cyc);
  run<F8>("8 FFMA2(3 regs)", out, cyc); run<F8_LOP8>("8 FFMA2 + 8 LOP3", out, cyc); run<F8_MOV8>("8 FFMA2 + 8 MOV", out, cyc);
  run<F8_LDS8>("8 FFMA2 + 8 LDS.64", out, cyc); run<F8_MUFU4>("8 FFMA2 + 4 MUFU", out, cyc); run<F8_STS2>("8 FFMA2 + 2 STS.64", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
