// C ABI of libpdeopt_b200 (see include/pdeopt_b200.h for the contract and the reference
// code each entry point replaces).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/pdeopt_b200.h"
#include "sifs128.cuh"

using namespace pdeopt;

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static pdeopt_status fail(pdeopt_status s, const std::string& msg) {
  g_err = msg;
  return s;
}
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct pdeopt_plan {
  pdeopt_plan_desc d;
  // scratch owned by the plan for the host-buffer entry point
  void* dev_scratch = nullptr;
  size_t dev_scratch_bytes = 0;
  float* park = nullptr;  // only used by PDEOPT_PARK_GLOBAL builds
  size_t park_bytes = 0;
  bool attr_set = false;
};

extern "C" int pdeopt_abi_version(void) { return PDEOPT_ABI_VERSION; }
extern "C" const char* pdeopt_last_error(void) { return g_err.c_str(); }
extern "C" int64_t pdeopt_launch_count(void) { return g_launches.load(); }

extern "C" pdeopt_status pdeopt_plan_create(const pdeopt_plan_desc* desc, pdeopt_plan** out) {
  if (!desc || !out) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (desc->kind != PDEOPT_CH2D && desc->kind != PDEOPT_AC2D)
    return fail(PDEOPT_ERR_UNSUPPORTED, "sifs: only CH2D / AC2D plans are implemented");
  if (desc->derivs != PDEOPT_DERIVS_FD) return fail(PDEOPT_ERR_UNSUPPORTED, "sifs: only derivs='fd' is implemented");
  if (desc->nx != 128 || desc->ny != 128) return fail(PDEOPT_ERR_UNSUPPORTED, "sifs: only 128x128 grids are implemented");
  if (!(desc->hx > 0) || !(desc->hy > 0)) return fail(PDEOPT_ERR_INVALID, "grid spacing must be positive");
  if (desc->mu_family < 0 || desc->mu_family > 3) return fail(PDEOPT_ERR_INVALID, "unknown mu family");
  if (desc->mob_family < 0 || desc->mob_family > 3) return fail(PDEOPT_ERR_INVALID, "unknown mobility family");
  if (desc->mu_ncoef < 0 || desc->mu_ncoef > PDEOPT_MAX_COEF || desc->mob_ncoef < 0 || desc->mob_ncoef > PDEOPT_MAX_COEF)
    return fail(PDEOPT_ERR_INVALID, "too many coefficients");
  pdeopt_plan* p = new (std::nothrow) pdeopt_plan();
  if (!p) return fail(PDEOPT_ERR_INVALID, "out of memory");
  p->d = *desc;
  *out = p;
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_plan_destroy(pdeopt_plan* plan) {
  if (!plan) return PDEOPT_OK;
  if (plan->dev_scratch) cudaFree(plan->dev_scratch);
  if (plan->park) cudaFree(plan->park);
  delete plan;
  return PDEOPT_OK;
}

extern "C" int64_t pdeopt_table_len(const pdeopt_plan* plan) {
  if (!plan) return 0;
  return (int64_t)(plan->d.nx / 2 + 1) * (plan->d.ny / 2 + 1);
}

template <int EQ, int MU, int MOB>
static cudaError_t launch(const SifsParams& p, int grid, cudaStream_t st) {
  auto kern = sifs128_kernel<EQ, MU, MOB>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SifsSmem));
    if (e != cudaSuccess) return e;
    attr = true;
  }
  kern<<<grid, kThreads, sizeof(SifsSmem), st>>>(p);
  return cudaGetLastError();
}

extern "C" pdeopt_status pdeopt_sifs_step_batched(pdeopt_plan* plan, const float* y0_dev, float* y1_dev, int32_t batch,
                                                  int32_t ksteps, const float* dt_host, const float* tables_dev,
                                                  int32_t ntab, const int32_t* tab_idx_host, const float* ctrl_dev,
                                                  uint8_t* obs_dev, float obs_lo, float obs_hi, float* reward_dev,
                                                  void* stream) {
  if (!plan || !y0_dev || !y1_dev || !dt_host || !tables_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0) return fail(PDEOPT_ERR_INVALID, "batch must be positive");
  if (ksteps <= 0 || ksteps > PDEOPT_MAX_FUSED_STEPS) return fail(PDEOPT_ERR_INVALID, "ksteps must be in [1, 64]");
  if (ntab <= 0 || ntab > PDEOPT_MAX_TABLES) return fail(PDEOPT_ERR_INVALID, "ntab must be 1 or 2");
  if (obs_dev && !(obs_hi > obs_lo)) return fail(PDEOPT_ERR_INVALID, "obs_hi must exceed obs_lo");
  const pdeopt_plan_desc& d = plan->d;
  SifsParams p;
  std::memset(&p, 0, sizeof(p));
  p.y0 = y0_dev;
  p.y1 = y1_dev;
  p.batch = batch;
  p.ksteps = ksteps;
  p.tables = tables_dev;
  p.ntab = ntab;
  p.ctrl = ctrl_dev;
  p.obs = obs_dev;
  p.obs_lo = obs_lo;
  p.obs_scale = obs_dev ? 1.0f / (obs_hi - obs_lo) : 0.f;
  p.reward = reward_dev;
  p.inv_hx = (float)(1.0 / d.hx);
  p.inv_hy = (float)(1.0 / d.hy);
  p.inv_hx2 = (float)(1.0 / (d.hx * d.hx));
  p.inv_hy2 = (float)(1.0 / (d.hy * d.hy));
  p.kappa = (float)d.kappa;
  p.lo_x = (float)d.lo_x;
  p.lo_y = (float)d.lo_y;
  p.hx = (float)d.hx;
  p.hy = (float)d.hy;
  p.pw.mu_family = d.mu_family;
  p.pw.mu_ncoef = d.mu_ncoef;
  p.pw.mob_family = d.mob_family;
  p.pw.mob_ncoef = d.mob_ncoef;
  for (int i = 0; i < PDEOPT_MAX_COEF; ++i) {
    p.pw.mu_coef[i] = (float)d.mu_coef[i];
    p.pw.mob_coef[i] = (float)d.mob_coef[i];
  }
  for (int k = 0; k < ksteps; ++k) {
    p.dt[k] = dt_host[k];
    const int t = tab_idx_host ? tab_idx_host[k] : 0;
    if (t < 0 || t >= ntab) return fail(PDEOPT_ERR_INVALID, "tab_idx out of range");
    p.tab[k] = (uint8_t)t;
  }
  const int grid = (batch + 1) / 2;
  cudaStream_t st = (cudaStream_t)stream;
#ifdef PDEOPT_PARK_GLOBAL
  {
    const size_t need = (size_t)grid * 32 * kThreads * sizeof(float2);
    if (plan->park_bytes < need) {
      if (plan->park) cudaFree(plan->park);
      plan->park = nullptr;
      plan->park_bytes = 0;
      CUDA_TRY(cudaMalloc((void**)&plan->park, need));
      plan->park_bytes = need;
    }
    p.park = plan->park;
  }
#endif
  cudaError_t e;
  if (d.kind == PDEOPT_AC2D) {
    e = launch<EQ_AC, MU_RUNTIME, MOB_RUNTIME>(p, grid, st);
  } else if (d.mu_family == PDEOPT_MU_LOG && d.mob_family == PDEOPT_MOB_DEGENERATE) {
    e = launch<EQ_CH, MU_LOG, MOB_DEGENERATE>(p, grid, st);
  } else if (d.mu_family == PDEOPT_MU_LOG && d.mob_family == PDEOPT_MOB_CONST) {
    e = launch<EQ_CH, MU_LOG, MOB_CONST>(p, grid, st);
  } else if (d.mu_family == PDEOPT_MU_DOUBLE_WELL && d.mob_family == PDEOPT_MOB_CONST) {
    e = launch<EQ_CH, MU_DOUBLE_WELL, MOB_CONST>(p, grid, st);
  } else {
    e = launch<EQ_CH, MU_RUNTIME, MOB_RUNTIME>(p, grid, st);
  }
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_sifs_step_batched_host(pdeopt_plan* plan, const float* y0_host, float* y1_host,
                                                       int32_t batch, int32_t ksteps, const float* dt_host,
                                                       const float* tables_host, int32_t ntab,
                                                       const int32_t* tab_idx_host, const float* ctrl_host,
                                                       uint8_t* obs_host, float obs_lo, float obs_hi,
                                                       float* reward_host, void* stream) {
  if (!plan || !y0_host || !y1_host || !dt_host || !tables_host) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0) return fail(PDEOPT_ERR_INVALID, "batch must be positive");
  if (ntab <= 0 || ntab > PDEOPT_MAX_TABLES) return fail(PDEOPT_ERR_INVALID, "ntab must be 1 or 2");
  const size_t npts = (size_t)plan->d.nx * plan->d.ny;
  const size_t y_bytes = (size_t)batch * npts * sizeof(float);
  const size_t tab_bytes = (size_t)ntab * pdeopt_table_len(plan) * sizeof(float);
  const size_t ctrl_bytes = (size_t)batch * PDEOPT_NCTRL * sizeof(float);
  const size_t obs_bytes = (size_t)batch * npts;
  const size_t rew_bytes = (size_t)batch * 2 * sizeof(float);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t need = al(y_bytes) + al(tab_bytes) + al(ctrl_bytes) + al(obs_bytes) + al(rew_bytes);
  if (plan->dev_scratch_bytes < need) {
    if (plan->dev_scratch) cudaFree(plan->dev_scratch);
    plan->dev_scratch = nullptr;
    plan->dev_scratch_bytes = 0;
    CUDA_TRY(cudaMalloc(&plan->dev_scratch, need));
    plan->dev_scratch_bytes = need;
  }
  char* base = (char*)plan->dev_scratch;
  float* y_dev = (float*)base;
  float* tab_dev = (float*)(base + al(y_bytes));
  float* ctrl_dev = (float*)(base + al(y_bytes) + al(tab_bytes));
  uint8_t* obs_dev = (uint8_t*)(base + al(y_bytes) + al(tab_bytes) + al(ctrl_bytes));
  float* rew_dev = (float*)(base + al(y_bytes) + al(tab_bytes) + al(ctrl_bytes) + al(obs_bytes));
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(y_dev, y0_host, y_bytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(tab_dev, tables_host, tab_bytes, cudaMemcpyHostToDevice, st));
  if (ctrl_host) CUDA_TRY(cudaMemcpyAsync(ctrl_dev, ctrl_host, ctrl_bytes, cudaMemcpyHostToDevice, st));
  pdeopt_status s = pdeopt_sifs_step_batched(plan, y_dev, y_dev, batch, ksteps, dt_host, tab_dev, ntab, tab_idx_host,
                                             ctrl_host ? ctrl_dev : nullptr, obs_host ? obs_dev : nullptr, obs_lo,
                                             obs_hi, reward_host ? rew_dev : nullptr, stream);
  if (s != PDEOPT_OK) return s;
  CUDA_TRY(cudaMemcpyAsync(y1_host, y_dev, y_bytes, cudaMemcpyDeviceToHost, st));
  if (obs_host) CUDA_TRY(cudaMemcpyAsync(obs_host, obs_dev, obs_bytes, cudaMemcpyDeviceToHost, st));
  if (reward_host) CUDA_TRY(cudaMemcpyAsync(reward_host, rew_dev, rew_bytes, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return PDEOPT_OK;
}
