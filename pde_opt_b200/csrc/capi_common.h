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

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)
