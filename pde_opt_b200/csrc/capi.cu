// C ABI of libpdeopt_b200 (see include/pdeopt_b200.h for the contract and the reference
// code each entry point replaces).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capi_common.h"
#include "sifs128.cuh"
#include "sifs128r.cuh"
#include "sifs_generic.cuh"
#include "sifs_small.cuh"
#include "ch_adjoint.cuh"
#include "ch_tangent.cuh"
#include "ch_given_mu.cuh"
#include "sbm.cuh"
#include "sifs128r_adj.cuh"
#include "fourier128.cuh"

using namespace pdeopt;

static thread_local std::string g_err;
namespace pdeopt_capi {
std::atomic<int64_t> g_launches{0};
pdeopt_status fail(pdeopt_status s, const std::string& msg) {
  g_err = msg;
  return s;
}
}  // namespace pdeopt_capi

struct pdeopt_plan {
  pdeopt_plan_desc d;
  // scratch owned by the plan for the host-buffer entry point
  void* dev_scratch = nullptr;
  size_t dev_scratch_bytes = 0;
  float* park = nullptr;  // only used by PDEOPT_PARK_GLOBAL builds
  size_t park_bytes = 0;
  bool attr_set = false;
  int32_t* flags = nullptr;  // caller-owned [batch] non-finite flags (pdeopt_plan_set_nonfinite_flags)
  float* traj = nullptr;     // set by pdeopt_sifs_rollout_fwd around its launch: checkpoint destination
  int save_every = 1;
  // derivs='fourier': wavenumber tables and the per-CTA scratch line
  float* kxy = nullptr;
  float2* fscratch = nullptr;
  size_t fscratch_bytes = 0;
  // pipeline of the host-buffer entry point
  static constexpr int kPipe = 3;
  cudaStream_t pipe[kPipe] = {};
  cudaEvent_t ev_start = nullptr, ev_done[kPipe] = {};
  bool pipe_ready = false;
};

extern "C" int pdeopt_abi_version(void) { return PDEOPT_ABI_VERSION; }
extern "C" const char* pdeopt_last_error(void) { return g_err.c_str(); }
extern "C" int64_t pdeopt_launch_count(void) { return g_launches.load(); }

extern "C" pdeopt_status pdeopt_plan_create(const pdeopt_plan_desc* desc, pdeopt_plan** out) {
  if (!desc || !out) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (desc->kind != PDEOPT_CH2D && desc->kind != PDEOPT_AC2D)
    return fail(PDEOPT_ERR_UNSUPPORTED, "sifs: only CH2D / AC2D plans are implemented");
  if (desc->derivs != PDEOPT_DERIVS_FD && desc->derivs != PDEOPT_DERIVS_FOURIER) return fail(PDEOPT_ERR_INVALID, "unknown derivs");
  if (desc->derivs == PDEOPT_DERIVS_FOURIER && !(desc->nx == 128 && desc->ny == 128))
    return fail(PDEOPT_ERR_UNSUPPORTED, "sifs: derivs='fourier' is implemented for 128x128 grids");
  {
    auto pow2 = [](int v) { return v >= 1 && (v & (v - 1)) == 0; };
    const bool tuned = desc->nx == 128 && desc->ny == 128;
    const bool generic = pow2(desc->nx) && pow2(desc->ny) && (int64_t)desc->nx * desc->ny <= kGenMaxPts &&
                         (int64_t)desc->nx * desc->ny >= 2;
    if (!tuned && !generic)
      return fail(PDEOPT_ERR_UNSUPPORTED,
                  "sifs: grids must be 128x128 (tuned kernel) or powers of two with nx*ny <= 8192 (generic kernel)");
  }
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
  if (plan->kxy) cudaFree(plan->kxy);
  if (plan->fscratch) cudaFree(plan->fscratch);
  if (plan->pipe_ready) {
    for (int i = 0; i < pdeopt_plan::kPipe; ++i) {
      cudaStreamDestroy(plan->pipe[i]);
      cudaEventDestroy(plan->ev_done[i]);
    }
    cudaEventDestroy(plan->ev_start);
  }
  delete plan;
  return PDEOPT_OK;
}

extern "C" int64_t pdeopt_table_len(const pdeopt_plan* plan) {
  if (!plan) return 0;
  return (int64_t)(plan->d.nx / 2 + 1) * (plan->d.ny / 2 + 1);
}

// The five instantiations of the fused 128x128 kernel are compiled in their own translation units
// (sifs128_inst_*.cu) so that the build parallelises; variant ids below.
enum : int { SIFS_V_AC_RT = 0, SIFS_V_CH_LOG_DEG = 1, SIFS_V_CH_LOG_CONST = 2, SIFS_V_CH_DW_CONST = 3, SIFS_V_CH_RT = 4 };
cudaError_t pdeopt_sifs128_launch_a(int variant, const pdeopt::SifsParams& p, int grid, cudaStream_t st);
cudaError_t pdeopt_sifs128_launch_b(int variant, const pdeopt::SifsParams& p, int grid, cudaStream_t st);
cudaError_t pdeopt_sifs128r_launch_a(int variant, const pdeopt::SifsParams& p, cudaStream_t st);
cudaError_t pdeopt_sifs128r_launch_b(int variant, const pdeopt::SifsParams& p, cudaStream_t st);
cudaError_t pdeopt_sifs128r_adj_launch(const pdeopt::rf::AdjParams& p, cudaStream_t st);
// PDEOPT_SIFS128_PAIR=1 selects the round-1 kernel (two environments per 512-thread CTA) for A/B
// measurements; the default is the one-field-per-CTA kernel (sifs128r.cuh).
static bool use_pair_kernel() {
  static const bool v = [] {
    const char* e = std::getenv("PDEOPT_SIFS128_PAIR");
    return e && e[0] == '1';
  }();
  return v;
}
static cudaError_t launch_variant(int variant, const SifsParams& p, int grid, cudaStream_t st) {
  const bool a = variant == SIFS_V_CH_LOG_DEG || variant == SIFS_V_CH_LOG_CONST;
  if (!use_pair_kernel()) return a ? pdeopt_sifs128r_launch_a(variant, p, st) : pdeopt_sifs128r_launch_b(variant, p, st);
  return a ? pdeopt_sifs128_launch_a(variant, p, grid, st) : pdeopt_sifs128_launch_b(variant, p, grid, st);
}

// 1 where the environment's field holds a NaN / Inf (the per-environment failure flag; the
// reference only has diffrax's `throw` switch, pde_model.py:131 / pde_env.py:293-303)
__global__ void nonfinite_flags_kernel(const float* __restrict__ y, long long n_per_env, int32_t* __restrict__ flags) {
  const float* ye = y + (size_t)blockIdx.x * n_per_env;
  int bad = 0;
  for (long long i = threadIdx.x; i < n_per_env; i += blockDim.x) bad |= !(fabsf(ye[i]) <= 3.0e38f);
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) flags[blockIdx.x] = bad ? 1 : 0;
}

extern "C" pdeopt_status pdeopt_nonfinite_flags(const float* y_dev, int32_t batch, int64_t n_per_env, int32_t* flags_dev,
                                                void* stream) {
  PdeoptDeviceGuard device_guard_(y_dev);
  if (!y_dev || !flags_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0 || n_per_env <= 0) return fail(PDEOPT_ERR_INVALID, "batch and n_per_env must be positive");
  nonfinite_flags_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(y_dev, (long long)n_per_env, flags_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_plan_set_nonfinite_flags(pdeopt_plan* plan, int32_t* flags_dev) {
  if (!plan) return fail(PDEOPT_ERR_INVALID, "null argument");
  plan->flags = flags_dev;
  return PDEOPT_OK;
}

// env_offset: index of the first environment of this launch within the caller's batch (the host-buffer
// entry point launches chunks of one batch on concurrent streams: plan-owned per-environment scratch
// and the flag array are addressed by the global environment index).
static pdeopt_status sifs_launch_impl(pdeopt_plan* plan, int mode, const float* f0_dev, const float* y0_dev, float* y1_dev,
                                 int32_t batch, int32_t ksteps, const float* dt_host, const float* symbol_dev,
                                 const float* ctrl_dev, uint8_t* obs_dev, float obs_lo, float obs_hi,
                                 float* reward_dev, void* stream, int32_t env_offset, int32_t total_batch);

static pdeopt_status sifs_launch(pdeopt_plan* plan, int mode, const float* f0_dev, const float* y0_dev, float* y1_dev,
                                 int32_t batch, int32_t ksteps, const float* dt_host, const float* symbol_dev,
                                 const float* ctrl_dev, uint8_t* obs_dev, float obs_lo, float obs_hi,
                                 float* reward_dev, void* stream, int32_t env_offset = 0, int32_t total_batch = 0) {
  pdeopt_status s = sifs_launch_impl(plan, mode, f0_dev, y0_dev, y1_dev, batch, ksteps, dt_host, symbol_dev, ctrl_dev,
                                     obs_dev, obs_lo, obs_hi, reward_dev, stream, env_offset,
                                     total_batch > 0 ? total_batch : batch);
  if (s != PDEOPT_OK || !plan->flags || mode == MODE_RHS_ONLY) return s;
  const bool fused128 = plan->d.nx == 128 && plan->d.ny == 128 && plan->d.derivs == PDEOPT_DERIVS_FD && !use_pair_kernel();
  if (fused128) return s;  // written by the kernel's epilogue
  return pdeopt_nonfinite_flags(y1_dev, batch, (int64_t)plan->d.nx * plan->d.ny, plan->flags + env_offset, stream);
}

static pdeopt_status sifs_launch_impl(pdeopt_plan* plan, int mode, const float* f0_dev, const float* y0_dev, float* y1_dev,
                                 int32_t batch, int32_t ksteps, const float* dt_host, const float* symbol_dev,
                                 const float* ctrl_dev, uint8_t* obs_dev, float obs_lo, float obs_hi,
                                 float* reward_dev, void* stream, int32_t env_offset, int32_t total_batch) {
  if (!plan || !y0_dev || !y1_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (mode != MODE_RHS_ONLY && (!dt_host || !symbol_dev)) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (mode == MODE_GIVEN_F && (!f0_dev || ksteps != 1)) return fail(PDEOPT_ERR_INVALID, "given-f mode needs f0 and ksteps == 1");
  if (mode == MODE_RHS_ONLY) ksteps = 1;
  if (batch <= 0) return fail(PDEOPT_ERR_INVALID, "batch must be positive");
  if (ksteps <= 0 || ksteps > PDEOPT_MAX_FUSED_STEPS) return fail(PDEOPT_ERR_INVALID, "ksteps must be in [1, 512]");
  if (obs_dev && !(obs_hi > obs_lo)) return fail(PDEOPT_ERR_INVALID, "obs_hi must exceed obs_lo");
  const pdeopt_plan_desc& d = plan->d;
  SifsParams p;
  std::memset(&p, 0, sizeof(p));
  p.y0 = y0_dev;
  p.y1 = y1_dev;
  p.batch = batch;
  p.ksteps = ksteps;
  p.symbol = (mode == MODE_RHS_ONLY) ? nullptr : symbol_dev;
  p.ctrl = ctrl_dev;
  p.obs = obs_dev;
  p.obs_lo = obs_lo;
  p.obs_scale = obs_dev ? 1.0f / (obs_hi - obs_lo) : 0.f;
  p.reward = reward_dev;
  p.mode = mode;
  p.f0 = f0_dev;
  p.nonfinite = (plan->flags && mode != MODE_RHS_ONLY) ? plan->flags + env_offset : nullptr;
  p.traj = plan->traj;
  p.save_every = plan->save_every > 0 ? plan->save_every : 1;
  p.inv_hx = (float)(1.0 / d.hx);
  p.inv_hy = (float)(1.0 / d.hy);
  p.inv_hx2 = (float)(1.0 / (d.hx * d.hx));
  p.inv_hy2 = (float)(1.0 / (d.hy * d.hy));
  p.kappa = (float)d.kappa;
  sifs_fill_rhs_consts(p, d.kind == PDEOPT_AC2D);
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
  for (int k = 0; k < ksteps && mode != MODE_RHS_ONLY; ++k) p.dt[k] = dt_host[k];
  const int grid = (batch + 1) / 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (d.derivs == PDEOPT_DERIVS_FOURIER) {
    // pseudo-spectral right-hand sides (cahn_hilliard.py:82-87, allen_cahn.py:74-79): one env per CTA
    if (mode == MODE_GIVEN_F) return fail(PDEOPT_ERR_INVALID, "given-f mode does not depend on derivs; use an fd plan");
    if (obs_dev || reward_dev) return fail(PDEOPT_ERR_UNSUPPORTED, "derivs='fourier': no observation / reward epilogue");
    if (!plan->kxy) {
      float h[2 * kN];
      const float two_pi = (float)6.283185307179586;  // complex64(2j) * complex64(pi)
      for (int i = 0; i < kN; ++i) {
        const int f = i < kN / 2 ? i : i - kN;  // numpy.fft.fftfreq ordering (Nyquist negative)
        h[i] = two_pi * (float)((double)f / ((double)kN * d.hx));
        h[kN + i] = two_pi * (float)((double)f / ((double)kN * d.hy));
      }
      CUDA_TRY(cudaMalloc((void**)&plan->kxy, sizeof(h)));
      CUDA_TRY(cudaMemcpy(plan->kxy, h, sizeof(h), cudaMemcpyHostToDevice));
    }
    // one scratch line per environment of the caller's WHOLE batch: chunks of one batch that run on
    // concurrent streams (pdeopt_sifs_step_batched_host) address disjoint lines
    const size_t need = (size_t)total_batch * 32 * kThreads * sizeof(float2);
    if (plan->fscratch_bytes < need) {
      if (plan->fscratch) cudaFree(plan->fscratch);
      plan->fscratch = nullptr;
      plan->fscratch_bytes = 0;
      CUDA_TRY(cudaMalloc((void**)&plan->fscratch, need));
      plan->fscratch_bytes = need;
    }
    FourierParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.y0 = y0_dev;
    fp.y1 = y1_dev;
    fp.batch = batch;
    fp.ksteps = ksteps;
    fp.mode = mode;
    fp.eq = d.kind == PDEOPT_AC2D ? EQ_AC : EQ_CH;
    fp.symbol = p.symbol;
    fp.kx = plan->kxy;
    fp.ky = plan->kxy + kN;
    fp.ctrl = ctrl_dev;
    fp.scratch = plan->fscratch + (size_t)env_offset * 32 * kThreads;
    fp.kappa = p.kappa;
    fp.lo_x = p.lo_x;
    fp.lo_y = p.lo_y;
    fp.hx = p.hx;
    fp.hy = p.hy;
    fp.pw = p.pw;
    for (int k = 0; k < ksteps && mode != MODE_RHS_ONLY; ++k) fp.dt[k] = dt_host[k];
    static bool fattr[kPdeoptMaxDevices] = {};
    if (pdeopt_first_use_on_device(fattr))
      CUDA_TRY(cudaFuncSetAttribute(fourier128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FourierSmem)));
    fourier128_kernel<<<batch, kThreads, sizeof(FourierSmem), st>>>(fp);
    cudaError_t fe = cudaGetLastError();
    if (fe != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(fe));
    g_launches.fetch_add(1);
    return PDEOPT_OK;
  }
  if (!(d.nx == 128 && d.ny == 128)) {
    GenParams gp;
    gp.s = p;
    gp.nx = d.nx;
    gp.ny = d.ny;
    gp.lognx = ilog2(d.nx);
    gp.logny = ilog2(d.ny);
    if (d.nx == d.ny && (d.nx == 64 || d.nx == 32)) {
      // tuned small-grid kernel (two register stages per axis, few barriers per step)
      cudaError_t se = cudaSuccess;
#define PDEOPT_SMALL_LAUNCH(NN, EE)                                                                                   \
  {                                                                                                                   \
    static bool sattr[kPdeoptMaxDevices] = {};                                                                        \
    if (pdeopt_first_use_on_device(sattr))                                                                            \
      se = cudaFuncSetAttribute(sifs_small_kernel<NN, EE>, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                                (int)sizeof(SmallSmem<NN>));                                                          \
    if (se == cudaSuccess) sifs_small_kernel<NN, EE><<<grid, kSmallThreads, sizeof(SmallSmem<NN>), st>>>(gp);         \
  }
      if (d.nx == 64) {
        if (d.kind == PDEOPT_AC2D) PDEOPT_SMALL_LAUNCH(64, EQ_AC) else PDEOPT_SMALL_LAUNCH(64, EQ_CH)
      } else {
        if (d.kind == PDEOPT_AC2D) PDEOPT_SMALL_LAUNCH(32, EQ_AC) else PDEOPT_SMALL_LAUNCH(32, EQ_CH)
      }
#undef PDEOPT_SMALL_LAUNCH
      if (se == cudaSuccess) se = cudaGetLastError();
      if (se != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(se));
      g_launches.fetch_add(1);
      return PDEOPT_OK;
    }
    const size_t smem = gen_smem_bytes(d.nx, d.ny);
    cudaError_t ge;
    if (d.kind == PDEOPT_AC2D) {
      static bool attr[kPdeoptMaxDevices] = {};
      if (pdeopt_first_use_on_device(attr)) {
        ge = cudaFuncSetAttribute(sifs_generic_kernel<EQ_AC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ge != cudaSuccess) return fail(PDEOPT_ERR_CUDA, cudaGetErrorString(ge));
      }
      sifs_generic_kernel<EQ_AC><<<grid, kGenThreads, smem, st>>>(gp);
    } else {
      static bool attr[kPdeoptMaxDevices] = {};
      if (pdeopt_first_use_on_device(attr)) {
        ge = cudaFuncSetAttribute(sifs_generic_kernel<EQ_CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ge != cudaSuccess) return fail(PDEOPT_ERR_CUDA, cudaGetErrorString(ge));
      }
      sifs_generic_kernel<EQ_CH><<<grid, kGenThreads, smem, st>>>(gp);
    }
    ge = cudaGetLastError();
    if (ge != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(ge));
    g_launches.fetch_add(1);
    return PDEOPT_OK;
  }
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
    e = launch_variant(SIFS_V_AC_RT, p, grid, st);
  } else if (d.mu_family == PDEOPT_MU_LOG && d.mob_family == PDEOPT_MOB_DEGENERATE) {
    e = launch_variant(SIFS_V_CH_LOG_DEG, p, grid, st);
  } else if (d.mu_family == PDEOPT_MU_LOG && d.mob_family == PDEOPT_MOB_CONST) {
    e = launch_variant(SIFS_V_CH_LOG_CONST, p, grid, st);
  } else if (d.mu_family == PDEOPT_MU_DOUBLE_WELL && d.mob_family == PDEOPT_MOB_CONST) {
    e = launch_variant(SIFS_V_CH_DW_CONST, p, grid, st);
  } else {
    e = launch_variant(SIFS_V_CH_RT, p, grid, st);
  }
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_sifs_step_batched(pdeopt_plan* plan, const float* y0_dev, float* y1_dev, int32_t batch,
                                                  int32_t ksteps, const float* dt_host, const float* symbol_dev,
                                                  const float* ctrl_dev, uint8_t* obs_dev, float obs_lo, float obs_hi,
                                                  float* reward_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(y0_dev);
  return sifs_launch(plan, MODE_FUSED, nullptr, y0_dev, y1_dev, batch, ksteps, dt_host, symbol_dev, ctrl_dev, obs_dev,
                     obs_lo, obs_hi, reward_dev, stream);
}

extern "C" pdeopt_status pdeopt_rhs_batched(pdeopt_plan* plan, const float* y_dev, float* f_dev, int32_t batch,
                                            const float* ctrl_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(y_dev);
  return sifs_launch(plan, MODE_RHS_ONLY, nullptr, y_dev, f_dev, batch, 1, nullptr, nullptr, ctrl_dev, nullptr, 0.f, 1.f,
                     nullptr, stream);
}

extern "C" pdeopt_status pdeopt_sifs_filter_batched(pdeopt_plan* plan, const float* y0_dev, const float* f0_dev,
                                                    float* y1_dev, int32_t batch, float dt, const float* symbol_dev,
                                                    void* stream) {
  PdeoptDeviceGuard device_guard_(y0_dev);
  return sifs_launch(plan, MODE_GIVEN_F, f0_dev, y0_dev, y1_dev, batch, 1, &dt, symbol_dev, nullptr, nullptr, 0.f, 1.f,
                     nullptr, stream);
}

static_assert(PDEOPT_ADJ_NCOEF == PDEOPT_MAX_COEF, "coefficient count mismatch");

extern "C" int64_t pdeopt_phasefield_adjoint_work_floats(const pdeopt_plan* plan, int32_t batch) {
  if (!plan || batch <= 0) return 0;
  return 6 * (int64_t)batch * plan->d.nx * plan->d.ny;
}

extern "C" pdeopt_status pdeopt_phasefield_adjoint_step(pdeopt_plan* plan, const float* u_dev, const float* lam1_dev,
                                                        float* lam0_dev, int32_t batch, float dt, const float* symbol_dev,
                                                        float* work_dev, double* gmu_dev, double* gmob_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(u_dev);
  if (!plan || !u_dev || !lam1_dev || !lam0_dev || !symbol_dev || !work_dev || !gmu_dev || !gmob_dev)
    return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0 || batch > 65535) return fail(PDEOPT_ERR_INVALID, "batch must be in [1, 65535]");
  const pdeopt_plan_desc& d = plan->d;
  if (d.derivs != PDEOPT_DERIVS_FD) return fail(PDEOPT_ERR_UNSUPPORTED, "adjoint: derivs='fd' only");
  const int64_t npts = (int64_t)d.nx * d.ny, n = npts * batch;
  float* zeros = work_dev;
  float* w = work_dev + n;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(zeros, 0, sizeof(float) * n, st));
  // w = dt G lam1 through the fused filter kernel (y1 = 0 + dt Re ifft(fft(lam1) / (1 + dt A sigma)))
  pdeopt_status s = sifs_launch(plan, MODE_GIVEN_F, lam1_dev, zeros, w, batch, 1, &dt, symbol_dev, nullptr, nullptr, 0.f, 1.f,
                                nullptr, stream);
  if (s != PDEOPT_OK) return s;
  ChAdjParams p;
  std::memset(&p, 0, sizeof(p));
  p.nx = d.nx; p.ny = d.ny; p.batch = batch; p.eq = d.kind == PDEOPT_AC2D ? 1 : 0;
  p.u = u_dev; p.w = w; p.lam1 = lam1_dev; p.lam0 = lam0_dev;
  p.mu = work_dev + 2 * n; p.dd = work_dev + 3 * n; p.mub = work_dev + 4 * n; p.db = work_dev + 5 * n;
  p.gmu = gmu_dev; p.gmob = gmob_dev;
  p.inv_hx = (float)(1.0 / d.hx); p.inv_hy = (float)(1.0 / d.hy);
  p.inv_hx2 = (float)(1.0 / (d.hx * d.hx)); p.inv_hy2 = (float)(1.0 / (d.hy * d.hy));
  p.kappa = (float)d.kappa;
  p.pw.mu_family = d.mu_family; p.pw.mu_ncoef = d.mu_ncoef; p.pw.mob_family = d.mob_family; p.pw.mob_ncoef = d.mob_ncoef;
  for (int i = 0; i < PDEOPT_MAX_COEF; ++i) { p.pw.mu_coef[i] = (float)d.mu_coef[i]; p.pw.mob_coef[i] = (float)d.mob_coef[i]; }
  dim3 grid((unsigned)((npts + 255) / 256), batch);
  ch_adj_mu_kernel<<<grid, 256, 0, st>>>(p);
  ch_adj_bar_kernel<<<grid, 256, 0, st>>>(p);
  ch_adj_out_kernel<<<grid, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("adjoint step: ") + cudaGetErrorString(e));
  g_launches.fetch_add(3);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_rhs_given_mu_batched(pdeopt_plan* plan, const float* u_dev, const float* muh_dev,
                                                     const float* mob_dev, float* f_dev, int32_t batch, float* work_dev,
                                                     void* stream) {
  PdeoptDeviceGuard device_guard_(u_dev);
  if (!plan || !u_dev || !muh_dev || !f_dev || !work_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0 || batch > 65535) return fail(PDEOPT_ERR_INVALID, "batch must be in [1, 65535]");
  const pdeopt_plan_desc& d = plan->d;
  const int64_t npts = (int64_t)d.nx * d.ny, n = npts * batch;
  GivenMuParams p;
  std::memset(&p, 0, sizeof(p));
  p.nx = d.nx; p.ny = d.ny; p.batch = batch; p.eq = d.kind == PDEOPT_AC2D ? 1 : 0;
  p.u = u_dev; p.muh = muh_dev; p.mob = mob_dev; p.mu = work_dev; p.dd = work_dev + n; p.f = f_dev;
  p.inv_hx = (float)(1.0 / d.hx); p.inv_hy = (float)(1.0 / d.hy);
  p.inv_hx2 = (float)(1.0 / (d.hx * d.hx)); p.inv_hy2 = (float)(1.0 / (d.hy * d.hy));
  p.kappa = (float)d.kappa;
  p.pw.mu_family = d.mu_family; p.pw.mu_ncoef = d.mu_ncoef; p.pw.mob_family = d.mob_family; p.pw.mob_ncoef = d.mob_ncoef;
  for (int i = 0; i < PDEOPT_MAX_COEF; ++i) { p.pw.mu_coef[i] = (float)d.mu_coef[i]; p.pw.mob_coef[i] = (float)d.mob_coef[i]; }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((npts + 255) / 256), batch);
  given_mu_pass1_kernel<<<grid, 256, 0, st>>>(p);
  int launches = 1;
  if (p.eq == 0) {
    given_mu_pass2_kernel<<<grid, 256, 0, st>>>(p);
    ++launches;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("rhs (given mu): ") + cudaGetErrorString(e));
  g_launches.fetch_add(launches);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_phasefield_adjoint_given_mu(pdeopt_plan* plan, const float* u_dev, const float* muh_dev,
                                                            const float* mob_dev, const float* lam1_dev, float* lam0_base_dev,
                                                            float* mubar_dev, float* dbar_dev, int32_t batch, float dt,
                                                            const float* symbol_dev, float* work_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(u_dev);
  if (!plan || !u_dev || !muh_dev || !lam1_dev || !lam0_base_dev || !mubar_dev || !dbar_dev || !symbol_dev || !work_dev)
    return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0 || batch > 65535) return fail(PDEOPT_ERR_INVALID, "batch must be in [1, 65535]");
  const pdeopt_plan_desc& d = plan->d;
  const int64_t npts = (int64_t)d.nx * d.ny, n = npts * batch;
  float* zeros = work_dev;
  float* w = work_dev + n;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(zeros, 0, sizeof(float) * n, st));
  int32_t* const flags = plan->flags;
  plan->flags = nullptr;
  pdeopt_status s = sifs_launch(plan, MODE_GIVEN_F, lam1_dev, zeros, w, batch, 1, &dt, symbol_dev, nullptr, nullptr, 0.f, 1.f,
                                nullptr, stream);  // w = dt G lam1
  plan->flags = flags;
  if (s != PDEOPT_OK) return s;
  GivenMuParams g;
  std::memset(&g, 0, sizeof(g));
  g.nx = d.nx; g.ny = d.ny; g.batch = batch; g.eq = d.kind == PDEOPT_AC2D ? 1 : 0; g.keep_mu = 1;
  g.u = u_dev; g.muh = muh_dev; g.mob = mob_dev; g.mu = work_dev + 2 * n; g.dd = work_dev + 3 * n; g.f = nullptr;
  g.inv_hx = (float)(1.0 / d.hx); g.inv_hy = (float)(1.0 / d.hy);
  g.inv_hx2 = (float)(1.0 / (d.hx * d.hx)); g.inv_hy2 = (float)(1.0 / (d.hy * d.hy));
  g.kappa = (float)d.kappa;
  g.pw.mu_family = d.mu_family; g.pw.mu_ncoef = d.mu_ncoef; g.pw.mob_family = d.mob_family; g.pw.mob_ncoef = d.mob_ncoef;
  for (int i = 0; i < PDEOPT_MAX_COEF; ++i) { g.pw.mu_coef[i] = (float)d.mu_coef[i]; g.pw.mob_coef[i] = (float)d.mob_coef[i]; }
  ChAdjParams p;
  std::memset(&p, 0, sizeof(p));
  p.nx = d.nx; p.ny = d.ny; p.batch = batch; p.eq = g.eq;
  p.u = u_dev; p.w = w; p.mu = g.mu; p.dd = g.dd; p.mub = mubar_dev; p.db = dbar_dev;
  p.inv_hx = g.inv_hx; p.inv_hy = g.inv_hy; p.inv_hx2 = g.inv_hx2; p.inv_hy2 = g.inv_hy2; p.kappa = g.kappa;
  dim3 grid((unsigned)((npts + 255) / 256), batch);
  given_mu_pass1_kernel<<<grid, 256, 0, st>>>(g);
  ch_adj_bar_kernel<<<grid, 256, 0, st>>>(p);
  sub_kappa_lap_kernel<<<grid, 256, 0, st>>>(lam1_dev, mubar_dev, lam0_base_dev, d.nx, d.ny, g.inv_hx2, g.inv_hy2, g.kappa);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("adjoint (given mu): ") + cudaGetErrorString(e));
  g_launches.fetch_add(3);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_sbm_rhs_batched(const pdeopt_sbm_desc* desc, const float* u_dev, const float* f_dev,
                                                const float* mu_dev, const float* mob_dev, const float* psi_dev,
                                                const float* ngp_dev, const float* side_dev, float cos_theta,
                                                float cos_pi_minus_theta, float flux, float* work_dev, float* out_dev,
                                                int32_t batch, void* stream) {
  PdeoptDeviceGuard device_guard_(u_dev);
  if (!desc || !u_dev || !f_dev || !mu_dev || !mob_dev || !psi_dev || !ngp_dev || !side_dev || !out_dev)
    return fail(PDEOPT_ERR_INVALID, "null argument");
  if (desc->nx <= 0 || desc->ny <= 0 || batch <= 0 || batch > 65535) return fail(PDEOPT_ERR_INVALID, "bad sizes");
  if (desc->kind != PDEOPT_CH2D && desc->kind != PDEOPT_AC2D) return fail(PDEOPT_ERR_INVALID, "kind must be CH2D or AC2D");
  if (desc->kind == PDEOPT_CH2D && !work_dev) return fail(PDEOPT_ERR_INVALID, "work_dev is required for Cahn-Hilliard");
  SbmParams p;
  std::memset(&p, 0, sizeof(p));
  p.nx = desc->nx; p.ny = desc->ny; p.batch = batch; p.eq = desc->kind == PDEOPT_AC2D ? 1 : 0;
  p.u = u_dev; p.fval = f_dev; p.muval = mu_dev; p.mob = mob_dev; p.psi = psi_dev; p.ngp = ngp_dev; p.lh = side_dev;
  p.inner = work_dev; p.out = out_dev;
  p.inv_hx = (float)(1.0 / desc->hx); p.inv_hy = (float)(1.0 / desc->hy);
  p.kappa = (float)desc->kappa; p.sqrt_kappa = (float)std::sqrt(desc->kappa);
  p.cos_a = cos_theta; p.cos_b = cos_pi_minus_theta; p.flux = flux;
  const int64_t npts = (int64_t)desc->nx * desc->ny;
  dim3 grid((unsigned)((npts + 255) / 256), batch);
  cudaStream_t st = (cudaStream_t)stream;
  sbm_pass1_kernel<<<grid, 256, 0, st>>>(p);
  int launches = 1;
  if (p.eq == 0) {
    sbm_pass2_kernel<<<grid, 256, 0, st>>>(p);
    ++launches;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("smoothed-boundary rhs: ") + cudaGetErrorString(e));
  g_launches.fetch_add(launches);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_sifs_rollout_fwd(pdeopt_plan* plan, const float* y0_dev, float* y1_dev, int32_t batch,
                                                 int32_t ksteps, const float* dt_host, const float* symbol_dev,
                                                 float* traj_dev, int32_t save_every, void* stream) {
  PdeoptDeviceGuard device_guard_(y0_dev);
  if (!plan || !y0_dev || !y1_dev || !dt_host || !symbol_dev || !traj_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0 || ksteps <= 0) return fail(PDEOPT_ERR_INVALID, "batch and ksteps must be positive");
  if (save_every <= 0 || save_every > PDEOPT_MAX_FUSED_STEPS) return fail(PDEOPT_ERR_INVALID, "save_every must be in [1, 512]");
  const pdeopt_plan_desc& d = plan->d;
  const size_t n = (size_t)batch * d.nx * d.ny;
  const bool fused128 = d.nx == 128 && d.ny == 128 && d.derivs == PDEOPT_DERIVS_FD && !use_pair_kernel();
  const float* cur = y0_dev;
  pdeopt_status s = PDEOPT_OK;
  if (fused128) {
    // the fused kernel writes the checkpoints itself: one launch per 512 steps
    const int chunk = (PDEOPT_MAX_FUSED_STEPS / save_every) * save_every;
    for (int k0 = 0; k0 < ksteps && s == PDEOPT_OK; k0 += chunk) {
      plan->traj = traj_dev + (size_t)(k0 / save_every) * n;
      plan->save_every = save_every;
      s = sifs_launch(plan, MODE_FUSED, nullptr, cur, y1_dev, batch, std::min(chunk, ksteps - k0), dt_host + k0, symbol_dev,
                      nullptr, nullptr, 0.f, 1.f, nullptr, stream);
      cur = y1_dev;
    }
    plan->traj = nullptr;
    plan->save_every = 1;
    return s;
  }
  for (int k0 = 0; k0 < ksteps; k0 += save_every) {
    CUDA_TRY(cudaMemcpyAsync(traj_dev + (size_t)(k0 / save_every) * n, cur, n * sizeof(float), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
    s = sifs_launch(plan, MODE_FUSED, nullptr, cur, y1_dev, batch, std::min((int)save_every, ksteps - k0), dt_host + k0,
                    symbol_dev, nullptr, nullptr, 0.f, 1.f, nullptr, stream);
    if (s != PDEOPT_OK) return s;
    cur = y1_dev;
  }
  return s;
}

extern "C" pdeopt_status pdeopt_sifs_rollout_bwd(pdeopt_plan* plan, const float* traj_dev, const float* lam1_dev,
                                                 float* lam0_dev, int32_t batch, int32_t ksteps, const float* dt_host,
                                                 const float* symbol_dev, float* work_dev, double* gmu_dev, double* gmob_dev,
                                                 void* stream) {
  PdeoptDeviceGuard device_guard_(traj_dev);
  if (!plan || !traj_dev || !lam1_dev || !lam0_dev || !dt_host || !symbol_dev || !gmu_dev || !gmob_dev)
    return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0 || ksteps <= 0) return fail(PDEOPT_ERR_INVALID, "batch and ksteps must be positive");
  const pdeopt_plan_desc& d = plan->d;
  if (d.derivs != PDEOPT_DERIVS_FD) return fail(PDEOPT_ERR_UNSUPPORTED, "adjoint: derivs='fd' only");
  const size_t n = (size_t)batch * d.nx * d.ny;
  if (!(d.nx == 128 && d.ny == 128)) {
    // other grids: the streaming adjoint step, in reverse order
    if (!work_dev) return fail(PDEOPT_ERR_INVALID, "work_dev is required for grids other than 128 x 128");
    const float* lam = lam1_dev;
    for (int k = ksteps - 1; k >= 0; --k) {
      pdeopt_status s = pdeopt_phasefield_adjoint_step(plan, traj_dev + (size_t)k * n, lam, lam0_dev, batch, dt_host[k], symbol_dev,
                                                       work_dev, gmu_dev, gmob_dev, stream);
      if (s != PDEOPT_OK) return s;
      lam = lam0_dev;
    }
    return PDEOPT_OK;
  }
  rf::AdjParams p;
  std::memset(&p, 0, sizeof(p));
  p.lam0 = lam0_dev; p.gmu = gmu_dev; p.gmob = gmob_dev; p.symbol = symbol_dev;
  p.batch = batch; p.eq = d.kind == PDEOPT_AC2D ? EQ_AC : EQ_CH;
  p.inv_hx = (float)(1.0 / d.hx); p.inv_hy = (float)(1.0 / d.hy);
  p.inv_hx2 = (float)(1.0 / (d.hx * d.hx)); p.inv_hy2 = (float)(1.0 / (d.hy * d.hy));
  p.kappa = (float)d.kappa;
  p.pw.mu_family = d.mu_family; p.pw.mu_ncoef = d.mu_ncoef; p.pw.mob_family = d.mob_family; p.pw.mob_ncoef = d.mob_ncoef;
  for (int i = 0; i < PDEOPT_MAX_COEF; ++i) { p.pw.mu_coef[i] = (float)d.mu_coef[i]; p.pw.mob_coef[i] = (float)d.mob_coef[i]; }
  const float* lam = lam1_dev;
  // launches of at most 512 steps, last steps first
  for (int k1 = ksteps; k1 > 0;) {
    const int k0 = std::max(0, k1 - PDEOPT_MAX_FUSED_STEPS);
    p.traj = traj_dev + (size_t)k0 * n;
    p.lam1 = lam;
    p.ksteps = k1 - k0;
    for (int k = 0; k < p.ksteps; ++k) p.dt[k] = dt_host[k0 + k];
    cudaError_t e = pdeopt_sifs128r_adj_launch(p, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("adjoint rollout: ") + cudaGetErrorString(e));
    g_launches.fetch_add(1);
    lam = lam0_dev;
    k1 = k0;
  }
  return PDEOPT_OK;
}

extern "C" int64_t pdeopt_phasefield_tangent_work_floats(const pdeopt_plan* plan, int32_t batch, int32_t ndir) {
  if (!plan || batch <= 0 || ndir <= 0) return 0;
  return (2 + 3 * (int64_t)ndir) * batch * plan->d.nx * plan->d.ny;
}

extern "C" pdeopt_status pdeopt_phasefield_tangent_steps(pdeopt_plan* plan, const float* traj_dev, float* v_dev, int32_t batch,
                                                         int32_t ndir, int32_t ksteps, const float* dt_host,
                                                         const float* dmu_dev, const float* dmob_dev, const float* symbol_dev,
                                                         float* work_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(v_dev);
  if (!plan || !traj_dev || !v_dev || !dt_host || !dmu_dev || !dmob_dev || !symbol_dev || !work_dev)
    return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0 || batch > 65535 || ndir <= 0 || ndir > 65535 || ksteps <= 0)
    return fail(PDEOPT_ERR_INVALID, "batch, ndir in [1, 65535] and ksteps > 0 required");
  const pdeopt_plan_desc& d = plan->d;
  if (d.derivs != PDEOPT_DERIVS_FD) return fail(PDEOPT_ERR_UNSUPPORTED, "tangent: derivs='fd' only");
  const int64_t npts = (int64_t)d.nx * d.ny, n = npts * batch;
  ChTanParams p;
  std::memset(&p, 0, sizeof(p));
  p.nx = d.nx; p.ny = d.ny; p.batch = batch; p.ndir = ndir; p.eq = d.kind == PDEOPT_AC2D ? 1 : 0;
  p.v = v_dev; p.dmu = dmu_dev; p.dmob = dmob_dev;
  p.mu = work_dev; p.dd = work_dev + n; p.mut = work_dev + 2 * n; p.ddt = p.mut + (int64_t)ndir * n; p.ft = p.ddt + (int64_t)ndir * n;
  p.inv_hx = (float)(1.0 / d.hx); p.inv_hy = (float)(1.0 / d.hy);
  p.inv_hx2 = (float)(1.0 / (d.hx * d.hx)); p.inv_hy2 = (float)(1.0 / (d.hy * d.hy));
  p.kappa = (float)d.kappa;
  p.pw.mu_family = d.mu_family; p.pw.mu_ncoef = d.mu_ncoef; p.pw.mob_family = d.mob_family; p.pw.mob_ncoef = d.mob_ncoef;
  for (int i = 0; i < PDEOPT_MAX_COEF; ++i) { p.pw.mu_coef[i] = (float)d.mu_coef[i]; p.pw.mob_coef[i] = (float)d.mob_coef[i]; }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((npts + 255) / 256), batch, ndir);
  int32_t* const flags = plan->flags;  // the caller's flag array is sized for `batch`, not for ndir * batch tangents
  plan->flags = nullptr;
  pdeopt_status s = PDEOPT_OK;
  for (int k = 0; k < ksteps && s == PDEOPT_OK; ++k) {
    p.u = traj_dev + (int64_t)k * n;
    ch_tan_mu_kernel<<<grid, 256, 0, st>>>(p);
    ch_tan_rhs_kernel<<<grid, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { s = fail(PDEOPT_ERR_CUDA, std::string("tangent step: ") + cudaGetErrorString(e)); break; }
    g_launches.fetch_add(2);
    // v <- v + dt G f~ : the filter of the forward step (solvers.py:62-63), ndir * batch fields at once
    const int32_t total = batch * ndir;
    for (int32_t b0 = 0; b0 < total && s == PDEOPT_OK; b0 += 32768) {
      const int32_t nb = std::min(32768, total - b0);
      s = sifs_launch(plan, MODE_GIVEN_F, p.ft + (int64_t)b0 * npts, v_dev + (int64_t)b0 * npts, v_dev + (int64_t)b0 * npts, nb, 1,
                      dt_host + k, symbol_dev, nullptr, nullptr, 0.f, 1.f, nullptr, stream);
    }
  }
  plan->flags = flags;
  return s;
}

extern "C" pdeopt_status pdeopt_sifs_step_batched_host(pdeopt_plan* plan, const float* y0_host, float* y1_host,
                                                       int32_t batch, int32_t ksteps, const float* dt_host,
                                                       const float* symbol_host, const float* ctrl_host,
                                                       uint8_t* obs_host, float obs_lo, float obs_hi,
                                                       float* reward_host, void* stream) {
  // Host buffers in / out.  The batch is cut into chunks that are pipelined over three internal
  // streams (H2D of chunk c+1, kernel of chunk c and D2H of chunk c-1 overlap; PCIe is full
  // duplex), ordered after the caller's stream and complete on return.
  if (!plan || !y0_host || !y1_host || !dt_host || !symbol_host) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0) return fail(PDEOPT_ERR_INVALID, "batch must be positive");
  const size_t npts = (size_t)plan->d.nx * plan->d.ny;
  const size_t y_bytes = (size_t)batch * npts * sizeof(float);
  const size_t tab_bytes = (size_t)pdeopt_table_len(plan) * sizeof(float);
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
  if (!plan->pipe_ready) {
    for (int i = 0; i < pdeopt_plan::kPipe; ++i) CUDA_TRY(cudaStreamCreateWithFlags(&plan->pipe[i], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&plan->ev_start, cudaEventDisableTiming));
    for (int i = 0; i < pdeopt_plan::kPipe; ++i) CUDA_TRY(cudaEventCreateWithFlags(&plan->ev_done[i], cudaEventDisableTiming));
    plan->pipe_ready = true;
  }
  char* base = (char*)plan->dev_scratch;
  float* y_dev = (float*)base;
  float* tab_dev = (float*)(base + al(y_bytes));
  float* ctrl_dev = (float*)(base + al(y_bytes) + al(tab_bytes));
  uint8_t* obs_dev = (uint8_t*)(base + al(y_bytes) + al(tab_bytes) + al(ctrl_bytes));
  float* rew_dev = (float*)(base + al(y_bytes) + al(tab_bytes) + al(ctrl_bytes) + al(obs_bytes));
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(tab_dev, symbol_host, tab_bytes, cudaMemcpyHostToDevice, st));
  if (ctrl_host) CUDA_TRY(cudaMemcpyAsync(ctrl_dev, ctrl_host, ctrl_bytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(plan->ev_start, st));
  // chunks of an even number of environments (a CTA owns a pair), at least ~2 waves of CTAs each
  int nchunk = batch >= 2048 ? 8 : (batch >= 512 ? 4 : 1);
  int per = ((batch + nchunk - 1) / nchunk + 1) & ~1;
  pdeopt_status status = PDEOPT_OK;
  for (int c = 0, b0 = 0; b0 < batch; ++c, b0 += per) {
    const int nb = (batch - b0 < per) ? batch - b0 : per;
    cudaStream_t ps = plan->pipe[c % pdeopt_plan::kPipe];
    CUDA_TRY(cudaStreamWaitEvent(ps, plan->ev_start, 0));
    float* yc = y_dev + (size_t)b0 * npts;
    CUDA_TRY(cudaMemcpyAsync(yc, y0_host + (size_t)b0 * npts, (size_t)nb * npts * sizeof(float), cudaMemcpyHostToDevice, ps));
    status = sifs_launch(plan, MODE_FUSED, nullptr, yc, yc, nb, ksteps, dt_host, tab_dev,
                         ctrl_host ? ctrl_dev + (size_t)b0 * PDEOPT_NCTRL : nullptr,
                         obs_host ? obs_dev + (size_t)b0 * npts : nullptr, obs_lo, obs_hi,
                         reward_host ? rew_dev + (size_t)b0 * 2 : nullptr, (void*)ps, b0, batch);
    if (status != PDEOPT_OK) break;
    CUDA_TRY(cudaMemcpyAsync(y1_host + (size_t)b0 * npts, yc, (size_t)nb * npts * sizeof(float), cudaMemcpyDeviceToHost, ps));
    if (obs_host) CUDA_TRY(cudaMemcpyAsync(obs_host + (size_t)b0 * npts, obs_dev + (size_t)b0 * npts, (size_t)nb * npts, cudaMemcpyDeviceToHost, ps));
    if (reward_host) CUDA_TRY(cudaMemcpyAsync(reward_host + (size_t)b0 * 2, rew_dev + (size_t)b0 * 2, (size_t)nb * 2 * sizeof(float), cudaMemcpyDeviceToHost, ps));
  }
  for (int i = 0; i < pdeopt_plan::kPipe; ++i) {
    cudaEventRecord(plan->ev_done[i], plan->pipe[i]);
    cudaStreamWaitEvent(st, plan->ev_done[i], 0);
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return status;
}

// ---- measured FP32 peak (FFMA chains), the denominator of the fused path's roofline ----------
__global__ void __launch_bounds__(512) fma_peak_kernel(float* out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0f + threadIdx.x * 1e-6f + i;
  const float b = 0.9999f, c = 1e-4f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" pdeopt_status pdeopt_measure_fp32_peak(double* tflops_out, void* stream) {
  if (!tflops_out) return fail(PDEOPT_ERR_INVALID, "null argument");
  int dev = 0, nsm = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int grid = nsm * 4, iters = 8192;
  float* out = nullptr;
  CUDA_TRY(cudaMalloc((void**)&out, sizeof(float) * grid * 512));
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, st));
    fma_peak_kernel<<<grid, 512, 0, st>>>(out, iters);
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 16.0 * iters * 512.0 * grid;
    const double tf = flop / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
    g_launches.fetch_add(1);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops_out = best;
  return PDEOPT_OK;
}
