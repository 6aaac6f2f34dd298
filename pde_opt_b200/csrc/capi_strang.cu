// C ABI: fused 128x128 Strang split-step (see include/pdeopt_b200.h).
#include "capi_common.h"
#include "strang128.cuh"
#include "vortex.cuh"
using namespace pdeopt;

extern "C" pdeopt_status pdeopt_strang_step_batched(const pdeopt_gpe_desc* desc, const float* y0_dev, float* y1_dev,
                                                    int32_t batch, int32_t ksteps, const float* dt_host,
                                                    const float* a_term_dev, float ts_re, float ts_im,
                                                    const float* ctrl_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(y0_dev);
  if (!desc || !y0_dev || !y1_dev || !dt_host) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (desc->nx != 128 || desc->ny != 128)
    return fail(PDEOPT_ERR_UNSUPPORTED, "strang: only 128x128 grids are implemented (256x256 needs the cluster kernel)");
  if (batch <= 0) return fail(PDEOPT_ERR_INVALID, "batch must be positive");
  if (ksteps <= 0 || ksteps > PDEOPT_MAX_FUSED_STEPS) return fail(PDEOPT_ERR_INVALID, "ksteps must be in [1, 512]");
  if (!(desc->hx > 0) || !(desc->hy > 0)) return fail(PDEOPT_ERR_INVALID, "grid spacing must be positive");
  StrangParams p;
  std::memset(&p, 0, sizeof(p));
  p.y0 = y0_dev;
  p.y1 = y1_dev;
  p.batch = batch;
  p.ksteps = ksteps;
  p.a_term = a_term_dev;
  p.ts_re = ts_re;
  p.ts_im = ts_im;
  p.dx = (float)desc->hx;
  p.k_int = (float)desc->k;
  p.e = (float)desc->e;
  p.trap = (float)desc->trap_factor;
  p.lo_x = (float)desc->lo_x;
  p.lo_y = (float)desc->lo_y;
  p.hx = (float)desc->hx;
  p.hy = (float)desc->hy;
  p.ctrl = ctrl_dev;
  for (int k = 0; k < ksteps; ++k) p.dt[k] = dt_host[k];
#ifdef PDEOPT_PARK_GLOBAL
  return fail(PDEOPT_ERR_UNSUPPORTED, "strang: PDEOPT_PARK_GLOBAL builds are not supported");
#endif
  static bool attr[kPdeoptMaxDevices] = {};
  if (pdeopt_first_use_on_device(attr))
    CUDA_TRY(cudaFuncSetAttribute(strang128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StrangSmem)));
  strang128_kernel<<<batch, kThreads, sizeof(StrangSmem), (cudaStream_t)stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}


extern "C" pdeopt_status pdeopt_gpe_detect_vortices(const float* psi_dev, int32_t batch, int32_t n0, int32_t n1,
                                                    float amp_thresh, float tol, int32_t* winding_dev,
                                                    int32_t* counts_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(psi_dev);
  if (!psi_dev || !counts_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0 || batch > 65535 || n0 < 2 || n1 < 2) return fail(PDEOPT_ERR_INVALID, "detect_vortices: bad sizes");
  VortexParams p;
  p.psi = reinterpret_cast<const float2*>(psi_dev);
  p.winding = winding_dev;
  p.counts = counts_dev;
  p.n0 = n0; p.n1 = n1; p.batch = batch;
  p.amp_thresh = amp_thresh; p.tol = tol;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(counts_dev, 0, sizeof(int32_t) * 3 * (size_t)batch, st));
  dim3 grid((n1 + 31) / 32, (n0 + 7) / 8, batch);
  vortex_kernel<<<grid, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("detect_vortices: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}
