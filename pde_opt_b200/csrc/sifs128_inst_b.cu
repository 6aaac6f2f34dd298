// Instantiations of the fused 128x128 SIFS kernel (double well and the runtime-switch variants); see capi.cu.
#include <cuda_runtime.h>

#include "sifs128.cuh"

using namespace pdeopt;

template <int EQ, int MU, int MOB>
static cudaError_t launch(const SifsParams& p, int grid, cudaStream_t st) {
  auto kern = sifs128_kernel<EQ, MU, MOB>;
  static bool attr[64] = {};  // per device
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !attr[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SifsSmem));
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  kern<<<grid, kThreads, sizeof(SifsSmem), st>>>(p);
  return cudaGetLastError();
}

cudaError_t pdeopt_sifs128_launch_b(int variant, const SifsParams& p, int grid, cudaStream_t st) {
  switch (variant) {
    case 0: return launch<EQ_AC, MU_RUNTIME, MOB_RUNTIME>(p, grid, st);
    case 3: return launch<EQ_CH, MU_DOUBLE_WELL, MOB_CONST>(p, grid, st);
    case 4: return launch<EQ_CH, MU_RUNTIME, MOB_RUNTIME>(p, grid, st);
    default: return cudaErrorInvalidValue;
  }
}
