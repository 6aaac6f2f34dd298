// Launch helper shared by the sifs128r_inst_*.cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

#include "sifs128r.cuh"

using namespace pdeopt;

// Per-device launch state: kernel attributes are per-device (and per-context) state, so they are
// set once for every device a kernel is launched on, not once per process.
constexpr int kMaxDevices = 64;

template <int EQ, int MU, int MOB>
static cudaError_t launch_r(const SifsParams& p, cudaStream_t st) {
  auto kern = rf::sifs128r_kernel<EQ, MU, MOB>;
  static bool attr[kMaxDevices] = {};
  static int sms[kMaxDevices] = {};
  // work counters of the dynamic environment distribution: a ring of slots per device so that launches
  // in flight on different streams use different counters; each is zeroed on its launch's stream
  constexpr int kSlots = 256;
  static int* counters[kMaxDevices] = {};
  static unsigned next_slot[kMaxDevices] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  if (!attr[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(rf::RSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void**)&counters[dev], kSlots * sizeof(int));
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  SifsParams q = p;
  q.work_counter = counters[dev] + (next_slot[dev]++ % kSlots);
  e = cudaMemsetAsync(q.work_counter, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  // PDEOPT_SIFS128R_ONE=1: one CTA per SM (experiment: how much do two co-resident CTAs overlap?)
  static const bool one = [] { const char* e = std::getenv("PDEOPT_SIFS128R_ONE"); return e && e[0] == '1'; }();
  const int slots = (one ? 1 : 2) * sms[dev];
  const int grid = p.batch < slots ? p.batch : slots;
  if (one) {
    static bool a2 = false;
    if (!a2) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024); a2 = true; }
  }
  kern<<<grid, rf::kThreadsR, one ? 120 * 1024 : sizeof(rf::RSmem), st>>>(q);
  return cudaGetLastError();
}
