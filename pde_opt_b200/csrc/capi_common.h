// Shared plumbing of the C ABI translation units (error string, launch counter).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/pdeopt_b200.h"

namespace pdeopt_capi {
pdeopt_status fail(pdeopt_status s, const std::string& msg);
extern std::atomic<int64_t> g_launches;
}  // namespace pdeopt_capi
using pdeopt_capi::fail;
using pdeopt_capi::g_launches;

// Kernel attributes (dynamic shared-memory limit, cluster flags) are per-device state: `flags` is a
// function-local static array indexed by the current device, so a process that uses several GPUs
// sets them once on each.
constexpr int kPdeoptMaxDevices = 64;
inline bool pdeopt_first_use_on_device(bool (&flags)[kPdeoptMaxDevices]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kPdeoptMaxDevices) return true;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

// Makes the device that owns `dev_ptr` current for the duration of one C-ABI call (the launch, the
// kernel attributes and any plan-owned scratch then all belong to the device of the data), and
// restores the caller's device on return.  Host pointers and NULL leave the current device alone.
struct PdeoptDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit PdeoptDeviceGuard(const void* dev_ptr) {
    cudaPointerAttributes a;
    if (dev_ptr && cudaPointerGetAttributes(&a, dev_ptr) == cudaSuccess &&
        (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged)) {
      if (cudaGetDevice(&prev) == cudaSuccess && prev != a.device && cudaSetDevice(a.device) == cudaSuccess) switched = true;
    } else {
      (void)cudaGetLastError();
    }
  }
  ~PdeoptDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  PdeoptDeviceGuard(const PdeoptDeviceGuard&) = delete;
  PdeoptDeviceGuard& operator=(const PdeoptDeviceGuard&) = delete;
};

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)
