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
    e = cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  const int grid = p.batch < sms[dev] ? p.batch : sms[dev];
  // the smallest instantiation that covers the coefficient counts of the plan's closures
  const int nmax = p.pw.mu_ncoef > p.pw.mob_ncoef ? p.pw.mu_ncoef : p.pw.mob_ncoef;
#define PDEOPT_ADJ_LAUNCH(EQ_, NC_)                                                                                   \
  do {                                                                                                                \
    auto kern = rf::sifs128r_adj_kernel<EQ_, NC_>;                                                                    \
    if (!attr[dev]) {                                                                                                 \
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(rf::ASmem));            \
      if (e != cudaSuccess) return e;                                                                                 \
    }                                                                                                                 \
    if (run) kern<<<grid, rf::kThreadsR, sizeof(rf::ASmem), st>>>(p);                                                 \
  } while (0)
  // first use on a device: set the attribute of every instantiation, then launch the selected one
  for (int pass = attr[dev] ? 1 : 0; pass < 2; ++pass) {
    const bool all = pass == 0;
    bool run;
    run = !all && p.eq != EQ_AC && nmax <= 4;             if (all || run) PDEOPT_ADJ_LAUNCH(EQ_CH, 4);
    run = !all && p.eq != EQ_AC && nmax > 4 && nmax <= 8; if (all || run) PDEOPT_ADJ_LAUNCH(EQ_CH, 8);
    run = !all && p.eq != EQ_AC && nmax > 8;              if (all || run) PDEOPT_ADJ_LAUNCH(EQ_CH, 16);
    run = !all && p.eq == EQ_AC && nmax <= 4;             if (all || run) PDEOPT_ADJ_LAUNCH(EQ_AC, 4);
    run = !all && p.eq == EQ_AC && nmax > 4 && nmax <= 8; if (all || run) PDEOPT_ADJ_LAUNCH(EQ_AC, 8);
    run = !all && p.eq == EQ_AC && nmax > 8;              if (all || run) PDEOPT_ADJ_LAUNCH(EQ_AC, 16);
    if (all) attr[dev] = true;
  }
#undef PDEOPT_ADJ_LAUNCH
  return cudaGetLastError();
}
