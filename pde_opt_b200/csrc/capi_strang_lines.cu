// C ABI: Strang split-step on the line-FFT engine and the cluster kernel (see include/pdeopt_b200.h).
#include "capi_lines_common.h"
#include "strang_lines.cuh"
#include "strang_cluster.cuh"
#include "strang_cluster_kin.cuh"

#include <cstdlib>

extern "C" int64_t pdeopt_strang_lines_work_floats(int32_t nx, int32_t ny, int32_t batch) {
  if (nx <= 0 || ny <= 0 || batch <= 0) return 0;
  return 2 * (int64_t)nx * ny * batch + 2 * (int64_t)nx * ny + 2 * (((int64_t)batch + 63) / 64) * 64;
}

extern "C" pdeopt_status pdeopt_strang_lines_step_batched(const pdeopt_gpe_desc* desc, const float* y0_dev, float* y1_dev,
                                                          int32_t batch, int32_t ksteps, const float* dt_host,
                                                          const float* a_term_full_dev, float ts_re, float ts_im,
                                                          const float* ctrl_dev, float* work_dev, void* stream) {
  return pdeopt_strang_lines_step_batched_light(desc, y0_dev, y1_dev, batch, ksteps, dt_host, a_term_full_dev, ts_re, ts_im,
                                                ctrl_dev, nullptr, 0, work_dev, stream);
}

extern "C" pdeopt_status pdeopt_strang_lines_step_batched_light(const pdeopt_gpe_desc* desc, const float* y0_dev,
                                                                float* y1_dev, int32_t batch, int32_t ksteps,
                                                                const float* dt_host, const float* a_term_full_dev,
                                                                float ts_re, float ts_im, const float* ctrl_dev,
                                                                const float* light_dev, int64_t light_env_stride,
                                                                float* work_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(y0_dev);
  if (!desc || !y0_dev || !y1_dev || !dt_host || !work_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (light_dev && light_env_stride != 0 && light_env_stride < (int64_t)desc->nx * desc->ny)
    return fail(PDEOPT_ERR_INVALID, "light_env_stride must be 0 (shared field) or >= nx*ny");
  const int nx = desc->nx, ny = desc->ny;
  if (!lf_size_ok(nx) || !lf_size_ok(ny) || nx < 32 || ny < 32)
    return fail(PDEOPT_ERR_UNSUPPORTED, "strang_lines: nx, ny must be powers of two in [32, 512]");
  if (batch <= 0) return fail(PDEOPT_ERR_INVALID, "batch must be positive");
  if (ksteps <= 0) return fail(PDEOPT_ERR_INVALID, "ksteps must be positive");
  if (!(desc->hx > 0) || !(desc->hy > 0)) return fail(PDEOPT_ERR_INVALID, "grid spacing must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (a_term_full_dev == nullptr && light_dev == nullptr && nx == kClN && ny == kClN) {
    // the equation as shipped on 256x256: cluster-of-4 kernel, state in registers for all K steps
    static bool cattr[kPdeoptMaxDevices] = {};
    if (pdeopt_first_use_on_device(cattr)) {
      cudaError_t ce = cudaFuncSetAttribute(strang_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
      (void)ce;
    }
    const float* src = y0_dev;
    for (int done = 0; done < ksteps;) {
      const int kk = ksteps - done < kMaxK ? ksteps - done : kMaxK;
      StrangClusterParams cp;
      std::memset(&cp, 0, sizeof(cp));
      cp.y0 = src; cp.y1 = y1_dev; cp.batch = batch; cp.ksteps = kk;
      cp.ts_re = ts_re; cp.ts_im = ts_im; cp.dx = (float)desc->hx;
      cp.k_int = (float)desc->k; cp.e = (float)desc->e; cp.trap = (float)desc->trap_factor;
      cp.lo_x = (float)desc->lo_x; cp.lo_y = (float)desc->lo_y; cp.hx = (float)desc->hx; cp.hy = (float)desc->hy;
      cp.ctrl = ctrl_dev;
      for (int k = 0; k < kk; ++k) cp.dt[k] = dt_host[done + k];
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(batch * kClCtas));
      cfg.blockDim = dim3(kClThreads);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = kClCtas;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t le = cudaLaunchKernelEx(&cfg, strang_cluster_kernel, cp);
      if (le != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("strang cluster launch: ") + cudaGetErrorString(le));
      g_launches.fetch_add(1);
      src = y1_dev;
      done += kk;
    }
    return PDEOPT_OK;
  }
  static const bool force_lines = [] { const char* e = std::getenv("PDEOPT_STRANG_LINES"); return e && e[0] == '1'; }();
  if (a_term_full_dev != nullptr && light_dev == nullptr && nx == cf::kN && ny == cf::kN && !force_lines) {
    // kinetic term on 256x256: cluster-of-4 kernel, the wavefunction stays inside the cluster for all K steps
    // (DSMEM transposes); one 512 KB kinetic table per distinct dt, built in the caller's scratch
    static bool kattr[kPdeoptMaxDevices] = {};
    if (pdeopt_first_use_on_device(kattr)) {
      CUDA_TRY(cudaFuncSetAttribute(cf::strang_cluster_kin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cf::KinSmem) + 2048));
      // two CTAs (of different clusters) per SM need the full shared-memory carve-out
      CUDA_TRY(cudaFuncSetAttribute(cf::strang_cluster_kin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    const int64_t np = (int64_t)cf::kN * cf::kN;
    const int max_tabs = batch < cf::kMaxTabs ? (batch < 1 ? 1 : batch) : cf::kMaxTabs;  // tables live in the W area
    float2* tabs = (float2*)work_dev;
    const float* src = y0_dev;
    int done = 0;
    while (done < ksteps) {
      cf::KinParams kp;
      std::memset(&kp, 0, sizeof(kp));
      float dts[cf::kMaxTabs];
      int ntab = 0, kk = 0;
      for (; done + kk < ksteps && kk < kMaxK; ++kk) {
        const float dt = dt_host[done + kk];
        int t = 0;
        while (t < ntab && dts[t] != dt) ++t;
        if (t == ntab) {
          if (ntab == max_tabs) break;
          dts[ntab++] = dt;
        }
        kp.dt[kk] = dt;
        kp.tab[kk] = (unsigned char)t;
      }
      for (int t = 0; t < ntab; ++t) {
        cf::strang_cluster_etab_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>((const float2*)a_term_full_dev, tabs + (size_t)t * np,
                                                                                   0.5f * dts[t] * ts_re, 0.5f * dts[t] * ts_im);
        g_launches.fetch_add(1);
      }
      kp.y0 = src; kp.y1 = y1_dev; kp.batch = batch; kp.ksteps = kk;
      kp.ts_re = ts_re; kp.ts_im = ts_im; kp.dx = (float)desc->hx;
      kp.k_int = (float)desc->k; kp.e = (float)desc->e; kp.trap = (float)desc->trap_factor;
      kp.lo_x = (float)desc->lo_x; kp.lo_y = (float)desc->lo_y; kp.hx = (float)desc->hx; kp.hy = (float)desc->hy;
      kp.ctrl = ctrl_dev; kp.etab = tabs;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(batch * cf::kCtas));
      cfg.blockDim = dim3(cf::kThreadsC);
      cfg.dynamicSmemBytes = sizeof(cf::KinSmem) + 2048;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cf::kCtas;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      static const bool dbg_occ = [] { const char* e = std::getenv("PDEOPT_DEBUG_OCC"); return e && e[0] == '1'; }();
      if (dbg_occ) {
        int ncl = -1, nblk = -1;
        cudaOccupancyMaxActiveClusters(&ncl, cf::strang_cluster_kin_kernel, &cfg);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, cf::strang_cluster_kin_kernel, cf::kThreadsC, cfg.dynamicSmemBytes);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, cf::strang_cluster_kin_kernel);
        fprintf(stderr, "strang_cluster_kin: max active clusters %d (x %d CTAs), CTAs per SM by resources %d, smem %zu B, regs %d, static smem %zu, local %zu, carveout %d, maxdyn %d\n", ncl, cf::kCtas,
                nblk, (size_t)cfg.dynamicSmemBytes, fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.preferredShmemCarveout, fa.maxDynamicSharedSizeBytes);
      }
      cudaError_t le = cudaLaunchKernelEx(&cfg, cf::strang_cluster_kin_kernel, kp);
      if (le != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("strang kinetic cluster launch: ") + cudaGetErrorString(le));
      g_launches.fetch_add(1);
      src = y1_dev;
      done += kk;
    }
    return PDEOPT_OK;
  }
  const int64_t npts = (int64_t)nx * ny, total = npts * batch;
  float2* W = (float2*)work_dev;
  float2* etab = W + total;
  float* norm = (float*)(etab + npts);
  const int64_t nstride = (((int64_t)batch + 63) / 64) * 64;
  GpeLinesConst c;
  c.nx = nx; c.ny = ny; c.log2nx = ilog2(nx);
  c.lo_x = (float)desc->lo_x; c.lo_y = (float)desc->lo_y; c.hx = (float)desc->hx; c.hy = (float)desc->hy;
  c.trap = (float)desc->trap_factor; c.e = (float)desc->e; c.k_int = (float)desc->k;
  c.ts_re = ts_re; c.ts_im = ts_im; c.ctrl = ctrl_dev;
  c.light = light_dev; c.light_env_stride = light_env_stride;
  const float dx2 = (float)desc->hx * (float)desc->hx;
  const LineGeom rows{(long long)batch * nx, 1, ny, 0, ny, 0, 1};
  const LineGeom cols{(long long)batch * ny, ny, npts, 1, nx, 0, ny};
  const float2* src = (const float2*)y0_dev;
  float2* dst = (float2*)y1_dev;
  float last_dt = 0.f;
  bool have_tab = false;
  cudaError_t e = cudaSuccess;
  for (int k = 0; k < ksteps && e == cudaSuccess; ++k) {
    const float dt = dt_host[k];
    if (a_term_full_dev == nullptr) {
      // one kernel per step; norms ping-pong so that step k can read the norm of step k-1 while
      // accumulating its own; the last norm is applied by the scale kernel after the loop
      float* nk = norm + (size_t)(k & 1) * nstride;
      const float* nprev = k > 0 ? norm + (size_t)((k - 1) & 1) * nstride : nullptr;
      e = cudaMemsetAsync(nk, 0, sizeof(float) * batch, st);
      if (e != cudaSuccess) break;
      const int bpe = (int)((npts + 16383) / 16384);
      // ping-pong between W and dst so that a launch never reads what it writes
      const float2* kin = k == 0 ? src : ((k & 1) ? W : dst);
      float2* kout = (k & 1) ? dst : W;
      strang_lines_potential_kernel<<<batch * bpe, 256, 0, st>>>(kin, kout, nprev, nk, c, dt, dx2, bpe);
      g_launches.fetch_add(1);
      if (k + 1 == ksteps) {
        strang_lines_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(kout, dst, nk, (int)npts, dx2, total);
        g_launches.fetch_add(1);
      }
      e = cudaGetLastError();
      continue;
    }
    e = cudaMemsetAsync(norm, 0, sizeof(float) * batch, st);
    if (e != cudaSuccess) break;
    {
      if (!have_tab || dt != last_dt) {
        strang_lines_etab_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>((const float2*)a_term_full_dev, etab, nx, ny,
                                                                               0.5f * dt * ts_re, 0.5f * dt * ts_im);
        have_tab = true;
        last_dt = dt;
        g_launches.fetch_add(1);
      }
      LineGeom ctab = cols;
      ctab.outer = 0;
      const LfMidCTab mid{etab, ctab};
      const LfMidCTabScaled mid_scaled{etab, norm, ctab, ilog2(ny), dx2};
      // 4 kernels per step (W holds the row-transformed state between steps):
      //   [rows fwd, first step only]  cols fwd*e*inv  rows inv*potential*fwd  cols fwd*e*scale*inv  rows inv -> y1 [-> fwd]
      if (k == 0) {
        e = lf_run<LF_FWD, true>(ny, rows.n_lines, LfLoadC{src, rows}, LfMidNone{}, LfStoreC{W, rows}, st);
        if (e != cudaSuccess) break;
        g_launches.fetch_add(1);
      }
      e = lf_run<LF_FWD_MUL_INV, false>(nx, cols.n_lines, LfLoadC{W, cols}, mid, LfStoreC{W, cols}, st);
      if (e != cudaSuccess) break;
      e = lf_run<LF_INV_MID_FWD, true>(ny, rows.n_lines, LfLoadC{W, rows}, LfMidPotentialIMF{src, norm, c, dt, 0.f}, LfStoreC{W, rows}, st);
      if (e != cudaSuccess) break;
      e = lf_run<LF_FWD_MUL_INV, false>(nx, cols.n_lines, LfLoadC{W, cols}, mid_scaled, LfStoreC{W, cols}, st);
      if (e != cudaSuccess) break;
      if (k + 1 < ksteps) {
        e = lf_run<LF_INV_MID_FWD, true>(ny, rows.n_lines, LfLoadC{W, rows}, LfMidStoreState{dst}, LfStoreC{W, rows}, st);
      } else {
        e = lf_run<LF_INV, true>(ny, rows.n_lines, LfLoadC{W, rows}, LfMidNone{}, LfStoreC{dst, rows}, st);
      }
      g_launches.fetch_add(4);
    }
    src = dst;
  }
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("strang_lines: ") + cudaGetErrorString(e));
  return PDEOPT_OK;
}

