// Instantiations of the fused 128x128 SIFS kernel (logarithmic potential, specialised packed code); see capi.cu.
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

cudaError_t pdeopt_sifs128_launch_a(int variant, const SifsParams& p, int grid, cudaStream_t st) {
  switch (variant) {
    case 1: return launch<EQ_CH, MU_LOG, MOB_DEGENERATE>(p, grid, st);
    case 2: return launch<EQ_CH, MU_LOG, MOB_CONST>(p, grid, st);
    default: return cudaErrorInvalidValue;
  }
}
