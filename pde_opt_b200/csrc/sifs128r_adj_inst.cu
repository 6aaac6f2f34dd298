// Instantiations and launcher of the fused K-step phase-field adjoint kernel (sifs128r_adj.cuh); see capi.cu.
#include "sifs128r_adj.cuh"

using namespace pdeopt;

cudaError_t pdeopt_sifs128r_adj_launch(const rf::AdjParams& p, cudaStream_t st) {
  constexpr int kMaxDev = 64;
  static bool attr[kMaxDev] = {};
  static int sms[kMaxDev] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDev) return cudaErrorInvalidDevice;
  if (!attr[dev]) {
    e = cudaFuncSetAttribute(rf::sifs128r_adj_kernel<EQ_CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(rf::ASmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(rf::sifs128r_adj_kernel<EQ_AC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(rf::ASmem));
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  const int grid = p.batch < sms[dev] ? p.batch : sms[dev];
  if (p.eq == EQ_AC)
    rf::sifs128r_adj_kernel<EQ_AC><<<grid, rf::kThreadsR, sizeof(rf::ASmem), st>>>(p);
  else
    rf::sifs128r_adj_kernel<EQ_CH><<<grid, rf::kThreadsR, sizeof(rf::ASmem), st>>>(p);
  return cudaGetLastError();
}
