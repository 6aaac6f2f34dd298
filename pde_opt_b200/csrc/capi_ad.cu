// C ABI: advection-diffusion rollout and adjoint (see include/pdeopt_b200.h).
#include "capi_common.h"
#include "ad128.cuh"
#include "ad_generic.cuh"
#include "ad_small.cuh"
using namespace pdeopt;

// ---- advection-diffusion rollout and its adjoint ------------------------------------------------
static pdeopt_status ad_fill(AdParams& p, const pdeopt_ad_desc* desc, int32_t batch, int32_t ksteps,
                             const float* dt_host, const float* tables_dev, const float* ctrl_dev, int32_t nseg,
                             int32_t hold, int32_t step0) {
  if (!desc || !dt_host || !tables_dev || !ctrl_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  {
    auto pow2 = [](int v) { return v >= 2 && (v & (v - 1)) == 0; };
    const bool tuned = desc->nx == 128 && desc->ny == 128;
    const bool generic = pow2(desc->nx) && pow2(desc->ny) && (int64_t)desc->nx * desc->ny <= kGenMaxPts;
    if (!tuned && !generic)
      return fail(PDEOPT_ERR_UNSUPPORTED,
                  "advection-diffusion: grids must be 128x128 (tuned kernels, with adjoint) or powers of two with "
                  "nx*ny <= 8192 (generic forward kernel)");
  }
  if (batch <= 0) return fail(PDEOPT_ERR_INVALID, "batch must be positive");
  if (ksteps <= 0 || ksteps > PDEOPT_MAX_FUSED_STEPS) return fail(PDEOPT_ERR_INVALID, "ksteps must be in [1, 512]");
  if (nseg <= 0 || hold <= 0 || step0 < 0) return fail(PDEOPT_ERR_INVALID, "bad control segmentation");
  if (!(desc->hx > 0) || !(desc->hy > 0)) return fail(PDEOPT_ERR_INVALID, "grid spacing must be positive");
  std::memset(&p, 0, sizeof(p));
  p.batch = batch;
  p.ksteps = ksteps;
  const int tl = (desc->nx / 2 + 1) * (desc->ny / 2 + 1);
  p.tabA = tables_dev;
  p.tabL = tables_dev + tl;
  p.kx = tables_dev + 2 * tl;
  p.ky = tables_dev + 2 * tl + desc->nx;
  p.ctrl = ctrl_dev;
  p.nseg = nseg;
  p.hold = hold;
  p.step0 = step0;
  p.lo_x = (float)desc->lo_x;
  p.lo_y = (float)desc->lo_y;
  p.hx = (float)desc->hx;
  p.hy = (float)desc->hy;
  for (int k = 0; k < ksteps; ++k) p.dt[k] = dt_host[k];
  return PDEOPT_OK;
}

extern "C" int64_t pdeopt_ad_tables_len(const pdeopt_ad_desc* desc) {
  if (!desc) return 0;
  return 2 * (int64_t)(desc->nx / 2 + 1) * (desc->ny / 2 + 1) + desc->nx + desc->ny;
}

extern "C" pdeopt_status pdeopt_ad_rollout_fwd(const pdeopt_ad_desc* desc, const float* y0_dev, float* y1_dev,
                                               int32_t batch, int32_t ksteps, const float* dt_host,
                                               const float* tables_dev, const float* ctrl_dev, int32_t nseg,
                                               int32_t hold, int32_t step0, float* traj_dev, int64_t traj_stride,
                                               void* stream) {
  PdeoptDeviceGuard device_guard_(y0_dev);
  if (!y0_dev || !y1_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
#ifdef PDEOPT_PARK_GLOBAL
  return fail(PDEOPT_ERR_UNSUPPORTED, "advection-diffusion: PDEOPT_PARK_GLOBAL builds are not supported");
#endif
  AdParams p;
  pdeopt_status s = ad_fill(p, desc, batch, ksteps, dt_host, tables_dev, ctrl_dev, nseg, hold, step0);
  if (s != PDEOPT_OK) return s;
  p.y0 = y0_dev;
  p.y1 = y1_dev;
  p.traj = traj_dev;
  p.traj_stride = traj_stride;
  if (!(desc->nx == 128 && desc->ny == 128)) {
    if (traj_dev) return fail(PDEOPT_ERR_UNSUPPORTED, "advection-diffusion: trajectory saving (adjoint) needs a 128x128 grid");
    AdGenParams gp;
    gp.a = p;
    gp.nx = desc->nx;
    gp.ny = desc->ny;
    gp.lognx = ilog2(desc->nx);
    gp.logny = ilog2(desc->ny);
    if (desc->nx == desc->ny && (desc->nx == 64 || desc->nx == 32)) {
      cudaError_t se;
      if (desc->nx == 64) {
        static bool a64[kPdeoptMaxDevices] = {};
        if (pdeopt_first_use_on_device(a64))
          CUDA_TRY(cudaFuncSetAttribute(ad_small_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AdSmallSmem<64>)));
        ad_small_fwd_kernel<64><<<(batch + 1) / 2, kSmallThreads, sizeof(AdSmallSmem<64>), (cudaStream_t)stream>>>(gp);
      } else {
        static bool a32[kPdeoptMaxDevices] = {};
        if (pdeopt_first_use_on_device(a32))
          CUDA_TRY(cudaFuncSetAttribute(ad_small_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AdSmallSmem<32>)));
        ad_small_fwd_kernel<32><<<(batch + 1) / 2, kSmallThreads, sizeof(AdSmallSmem<32>), (cudaStream_t)stream>>>(gp);
      }
      se = cudaGetLastError();
      if (se != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(se));
      g_launches.fetch_add(1);
      return PDEOPT_OK;
    }
    static bool gattr[kPdeoptMaxDevices] = {};
    if (pdeopt_first_use_on_device(gattr))
      CUDA_TRY(cudaFuncSetAttribute(ad_generic_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    ad_generic_fwd_kernel<<<(batch + 1) / 2, kGenThreads, ad_gen_smem_bytes(desc->nx, desc->ny), (cudaStream_t)stream>>>(gp);
    cudaError_t ge = cudaGetLastError();
    if (ge != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(ge));
    g_launches.fetch_add(1);
    return PDEOPT_OK;
  }
  static bool attr[kPdeoptMaxDevices] = {};
  if (pdeopt_first_use_on_device(attr))
    CUDA_TRY(cudaFuncSetAttribute(ad128_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AdSmem)));
  ad128_fwd_kernel<<<(batch + 1) / 2, kThreads, sizeof(AdSmem), (cudaStream_t)stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_ad_rollout_bwd(const pdeopt_ad_desc* desc, const float* traj_dev, int64_t traj_stride,
                                               const float* lam1_dev, float* lam0_dev, int32_t batch, int32_t ksteps,
                                               const float* dt_host, const float* tables_dev, const float* ctrl_dev,
                                               int32_t nseg, int32_t hold, int32_t step0, float* gctrl_dev,
                                               void* stream) {
  PdeoptDeviceGuard device_guard_(lam1_dev);
  if (!traj_dev || !lam1_dev || !lam0_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (desc && !(desc->nx == 128 && desc->ny == 128))
    return fail(PDEOPT_ERR_UNSUPPORTED, "advection-diffusion: the adjoint kernel is implemented for 128x128 grids");
#ifdef PDEOPT_PARK_GLOBAL
  return fail(PDEOPT_ERR_UNSUPPORTED, "advection-diffusion: PDEOPT_PARK_GLOBAL builds are not supported");
#endif
  AdParams p;
  pdeopt_status s = ad_fill(p, desc, batch, ksteps, dt_host, tables_dev, ctrl_dev, nseg, hold, step0);
  if (s != PDEOPT_OK) return s;
  p.y0 = lam1_dev;
  p.y1 = lam0_dev;
  p.traj_in = traj_dev;
  p.traj_stride = traj_stride;
  p.gctrl = gctrl_dev;
  static bool attr[kPdeoptMaxDevices] = {};
  if (pdeopt_first_use_on_device(attr))
    CUDA_TRY(cudaFuncSetAttribute(ad128_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AdSmem)));
  ad128_bwd_kernel<<<(batch + 1) / 2, kThreads, sizeof(AdSmem), (cudaStream_t)stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

