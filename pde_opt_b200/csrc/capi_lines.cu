// C ABI: line-FFT engine, 3-D Cahn-Hilliard, Strang on large grids (see include/pdeopt_b200.h).
#include <algorithm>
#include <cstdlib>

#include "capi_common.h"
#include "capi_lines_common.h"
#include "ch3d.cuh"
#include "ch_adjoint.cuh"
using namespace pdeopt;

// ---- line-FFT engine, 3-D Cahn-Hilliard, Strang on large grids ---------------------------------
extern "C" int32_t pdeopt_fft_pos_to_freq(int32_t n, int32_t pos) {
  if (!lf_size_ok(n) || pos < 0 || pos >= n) return -1;
  return line_pos_to_freq(n, pos);
}

extern "C" pdeopt_status pdeopt_fft_lines(const void* in_dev, void* out_dev, int32_t n, const pdeopt_line_geom* gin,
                                          const pdeopt_line_geom* gout, int32_t inverse, int32_t in_real, float scale,
                                          void* stream) {
  PdeoptDeviceGuard device_guard_(in_dev);
  if (!in_dev || !out_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (!lf_size_ok(n)) return fail(PDEOPT_ERR_UNSUPPORTED, "fft_lines: n must be a power of two in [8, 512]");
  if (!geom_ok(gin, n) || !geom_ok(gout, n) || gin->n_lines != gout->n_lines) return fail(PDEOPT_ERR_INVALID, "bad line geometry");
  const LineGeom gi = to_geom(gin), go = to_geom(gout);
  const bool contig = gi.lo == 1 && go.lo == 1 && gi.chunk == n && go.chunk == n;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  LfStoreCScaled sto{(float2*)out_dev, go, scale};
#define PDEOPT_RUN(MODE, LOADER)                                                                   \
  e = contig ? lf_run<MODE, true>(n, gi.n_lines, LOADER, LfMidNone{}, sto, st)                     \
             : lf_run<MODE, false>(n, gi.n_lines, LOADER, LfMidNone{}, sto, st)
  if (in_real) {
    LfLoadReal ld{(const float*)in_dev, gi};
    if (inverse) { PDEOPT_RUN(LF_INV, ld); } else { PDEOPT_RUN(LF_FWD, ld); }
  } else {
    LfLoadC ld{(const float2*)in_dev, gi};
    if (inverse) { PDEOPT_RUN(LF_INV, ld); } else { PDEOPT_RUN(LF_FWD, ld); }
  }
#undef PDEOPT_RUN
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("fft_lines: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_fft_lines_imex(const void* in_dev, void* out_dev, int32_t n, const pdeopt_line_geom* g,
                                               const float* sym_dev, const pdeopt_line_geom* gsym, float dt, float scale,
                                               void* stream) {
  PdeoptDeviceGuard device_guard_(in_dev);
  if (!in_dev || !out_dev || !sym_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (!lf_size_ok(n)) return fail(PDEOPT_ERR_UNSUPPORTED, "fft_lines: n must be a power of two in [8, 512]");
  if (!geom_ok(g, n) || !geom_ok(gsym, n)) return fail(PDEOPT_ERR_INVALID, "bad line geometry");
  const LineGeom gg = to_geom(g);
  LfLoadC ld{(const float2*)in_dev, gg};
  LfMidImex mid{sym_dev, to_geom(gsym), dt, scale};
  LfStoreC sto{(float2*)out_dev, gg};
  cudaError_t e = (gg.lo == 1 && gg.chunk == n) ? lf_run<LF_FWD_MUL_INV, true>(n, gg.n_lines, ld, mid, sto, (cudaStream_t)stream)
                                                : lf_run<LF_FWD_MUL_INV, false>(n, gg.n_lines, ld, mid, sto, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("fft_lines_imex: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_fft_lines_to_peers(const void* in_dev, int32_t n, const pdeopt_line_geom* gin,
                                                   void* const* peer_ptrs_host, int32_t n_peers,
                                                   const pdeopt_line_geom* gout, int64_t src_off, const float* sym_dev,
                                                   const pdeopt_line_geom* gsym, float dt, float scale, void* stream) {
  PdeoptDeviceGuard device_guard_(in_dev);
  if (!in_dev || !peer_ptrs_host) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (!lf_size_ok(n)) return fail(PDEOPT_ERR_UNSUPPORTED, "fft_lines: n must be a power of two in [8, 512]");
  if (n_peers < 1 || n_peers > 8) return fail(PDEOPT_ERR_INVALID, "1..8 peers");
  if (!geom_ok(gin, n) || !geom_ok(gout, n) || gin->n_lines != gout->n_lines) return fail(PDEOPT_ERR_INVALID, "bad line geometry");
  if (gout->chunk * n_peers != n || gout->hi != 0) return fail(PDEOPT_ERR_INVALID, "peer scatter: chunk * n_peers must equal n and hi must be 0");
  if (sym_dev && !geom_ok(gsym, n)) return fail(PDEOPT_ERR_INVALID, "bad symbol geometry");
  const LineGeom gi = to_geom(gin), go = to_geom(gout);
  if (gi.lo == 1 || go.lo == 1) return fail(PDEOPT_ERR_UNSUPPORTED, "peer scatter is for strided lines");
  LfStorePeers sto;
  for (int i = 0; i < 8; ++i) sto.peers[i] = (float2*)(i < n_peers ? peer_ptrs_host[i] : nullptr);
  sto.g = go;
  sto.shift = ilog2(gout->chunk);
  sto.src_off = src_off;
  LfLoadC ld{(const float2*)in_dev, gi};
  cudaError_t e;
  if (sym_dev) {
    LfMidImex mid{sym_dev, to_geom(gsym), dt, scale};
    e = lf_run<LF_FWD_MUL_INV, false>(n, gi.n_lines, ld, mid, sto, (cudaStream_t)stream);
  } else {
    e = lf_run<LF_FWD, false>(n, gi.n_lines, ld, LfMidNone{}, sto, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("fft_lines_to_peers: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

// Slab transpose as a dedicated push: block p of a contiguous send buffer goes to peer p's buffer (P2P stores over
// NVLink, 16 bytes per thread, fully coalesced).  Measured on this pool (tools/p2p_store_bench.cu): plain SM stores into
// peer memory reach 650-700 GB/s, while the same bytes stored from inside the last stage of a line-FFT kernel
// (pdeopt_fft_lines_to_peers) reach 270-390 GB/s — the remote stores back up the kernel's load / compute phases.
struct PushParams {
  const float4* src;
  float4* dst[8];
  long long src_block_vec4;  // distance between the blocks of consecutive peers in src (float4 elements)
  long long dst_off_vec4;    // offset inside every peer buffer
  long long run_vec4;        // contiguous run per row
  long long row_stride_vec4; // distance between rows (same in src and dst)
  int n_rows;
  int n_peers, first;        // block order starts at `first` so that the ranks do not all hit the same peer at once
};

static __global__ void __launch_bounds__(256) push_blocks_kernel(const __grid_constant__ PushParams p) {
  const long long per_block = p.run_vec4 * p.n_rows;
  const long long total = per_block * p.n_peers;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i / per_block);
    const long long w = i - (long long)k * per_block;
    const long long row = w / p.run_vec4, col = w - row * p.run_vec4;
    const int peer = (p.first + k) % p.n_peers;
    const long long o = row * p.row_stride_vec4 + col;
    p.dst[peer][p.dst_off_vec4 + o] = __ldg(p.src + (long long)peer * p.src_block_vec4 + o);
  }
}

static pdeopt_status push_launch(const void* src_dev, void* const* peer_ptrs_host, int32_t n_peers, int64_t src_block_bytes,
                                 int64_t dst_off_bytes, int32_t n_rows, int64_t run_bytes, int64_t row_stride_bytes,
                                 int32_t first_peer, void* stream) {
  PdeoptDeviceGuard device_guard_(src_dev);
  if (!src_dev || !peer_ptrs_host) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (n_peers < 1 || n_peers > 8) return fail(PDEOPT_ERR_INVALID, "1..8 peers");
  if (n_rows < 1 || run_bytes <= 0 || ((run_bytes | dst_off_bytes | src_block_bytes | row_stride_bytes) & 15) || dst_off_bytes < 0 ||
      src_block_bytes < 0 || row_stride_bytes < 0 || ((uintptr_t)src_dev & 15))
    return fail(PDEOPT_ERR_INVALID, "push: sizes, strides, offsets and the source pointer must be multiples of 16 bytes");
  PushParams p;
  std::memset(&p, 0, sizeof(p));
  p.src = (const float4*)src_dev;
  for (int i = 0; i < n_peers; ++i) {
    if (!peer_ptrs_host[i]) return fail(PDEOPT_ERR_INVALID, "null peer pointer");
    p.dst[i] = (float4*)peer_ptrs_host[i];
  }
  p.src_block_vec4 = src_block_bytes / 16;
  p.dst_off_vec4 = dst_off_bytes / 16;
  p.run_vec4 = run_bytes / 16;
  p.row_stride_vec4 = row_stride_bytes / 16;
  p.n_rows = n_rows;
  p.n_peers = n_peers;
  p.first = ((first_peer % n_peers) + n_peers) % n_peers;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long total = p.run_vec4 * n_rows * n_peers;
  const long long want = (total + 255) / 256;
  const int grid = (int)std::min<long long>((long long)sms * 8, std::max<long long>(want, 1));
  push_blocks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("push_blocks_to_peers: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_push_blocks_to_peers(const void* src_dev, void* const* peer_ptrs_host, int32_t n_peers,
                                                     int64_t block_bytes, int64_t dst_off_bytes, int32_t first_peer,
                                                     void* stream) {
  return push_launch(src_dev, peer_ptrs_host, n_peers, block_bytes, dst_off_bytes, 1, block_bytes, 0, first_peer, stream);
}

extern "C" pdeopt_status pdeopt_push_rows_to_peers(const void* src_dev, void* const* peer_ptrs_host, int32_t n_peers,
                                                   int64_t src_block_bytes, int64_t dst_off_bytes, int32_t n_rows,
                                                   int64_t run_bytes, int64_t row_stride_bytes, int32_t first_peer,
                                                   void* stream) {
  return push_launch(src_dev, peer_ptrs_host, n_peers, src_block_bytes, dst_off_bytes, n_rows, run_bytes, row_stride_bytes,
                     first_peer, stream);
}

extern "C" pdeopt_status pdeopt_fft_lines_inv_update(const void* spec_dev, int32_t n, const pdeopt_line_geom* gin,
                                                     const float* y0_dev, float* y1_dev, const pdeopt_line_geom* gout,
                                                     float dt, void* stream) {
  PdeoptDeviceGuard device_guard_(spec_dev);
  if (!spec_dev || !y0_dev || !y1_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (!lf_size_ok(n)) return fail(PDEOPT_ERR_UNSUPPORTED, "fft_lines: n must be a power of two in [8, 512]");
  if (!geom_ok(gin, n) || !geom_ok(gout, n) || gin->n_lines != gout->n_lines) return fail(PDEOPT_ERR_INVALID, "bad line geometry");
  const LineGeom gi = to_geom(gin), go = to_geom(gout);
  LfLoadC ld{(const float2*)spec_dev, gi};
  LfStoreUpdate sto{y0_dev, y1_dev, go, dt};
  const bool contig = gi.lo == 1 && go.lo == 1 && gi.chunk == n && go.chunk == n;
  cudaError_t e = contig ? lf_run<LF_INV, true>(n, gi.n_lines, ld, LfMidNone{}, sto, (cudaStream_t)stream)
                         : lf_run<LF_INV, false>(n, gi.n_lines, ld, LfMidNone{}, sto, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("fft_lines_inv_update: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_fft_lines_r2c(const float* in_dev, void* out_dev, int32_t n, int64_t n_lines, void* stream) {
  PdeoptDeviceGuard device_guard_(in_dev);
  if (!in_dev || !out_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (!lf_size_ok(n)) return fail(PDEOPT_ERR_UNSUPPORTED, "fft_lines: n must be a power of two in [8, 512]");
  if (n_lines <= 0 || (n_lines & 1)) return fail(PDEOPT_ERR_INVALID, "r2c needs a positive even number of lines");
  LfIoR2C io{in_dev, (float2*)out_dev, n, n / 2 + 1};
  cudaError_t e = cudaSuccess;
  auto run = [&]() -> cudaError_t {
    PDEOPT_LF_DISPATCH(n, return (lf_launch_real<LFN, false>(n_lines / 2, io, (cudaStream_t)stream)));
    return cudaSuccess;
  };
  e = run();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("fft_lines_r2c: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_fft_lines_c2r_update(const void* half_dev, int32_t n, int64_t n_lines, const float* y0_dev,
                                                     float* y1_dev, float dt, void* stream) {
  PdeoptDeviceGuard device_guard_(half_dev);
  if (!half_dev || !y0_dev || !y1_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (!lf_size_ok(n)) return fail(PDEOPT_ERR_UNSUPPORTED, "fft_lines: n must be a power of two in [8, 512]");
  if (n_lines <= 0 || (n_lines & 1)) return fail(PDEOPT_ERR_INVALID, "c2r needs a positive even number of lines");
  LfIoC2RUpdate io{(const float2*)half_dev, y0_dev, y1_dev, n, n / 2 + 1, dt};
  auto run = [&]() -> cudaError_t {
    PDEOPT_LF_DISPATCH(n, return (lf_launch_real<LFN, true>(n_lines / 2, io, (cudaStream_t)stream)));
    return cudaSuccess;
  };
  cudaError_t e = run();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("fft_lines_c2r_update: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return PDEOPT_OK;
}

static pdeopt_status ch3d_check(const pdeopt_ch3d_desc* d, int32_t batch) {
  if (!d) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (batch <= 0) return fail(PDEOPT_ERR_INVALID, "batch must be positive");
  if (d->nx < 1 || d->ny < 2 || d->nz < 2) return fail(PDEOPT_ERR_INVALID, "bad grid");
  if ((int64_t)batch * (d->nx + 2) > 65535) return fail(PDEOPT_ERR_UNSUPPORTED, "ch3d: batch * (nx + 2) must be <= 65535");
  if (d->ny > 65535) return fail(PDEOPT_ERR_UNSUPPORTED, "ch3d: ny must be <= 65535");
  if (!(d->hx > 0) || !(d->hy > 0) || !(d->hz > 0)) return fail(PDEOPT_ERR_INVALID, "grid spacing must be positive");
  if (d->mu_family < 0 || d->mu_family > 3 || d->mob_family < 0 || d->mob_family > 3) return fail(PDEOPT_ERR_INVALID, "unknown closure family");
  if (d->mu_ncoef < 0 || d->mu_ncoef > PDEOPT_MAX_COEF || d->mob_ncoef < 0 || d->mob_ncoef > PDEOPT_MAX_COEF)
    return fail(PDEOPT_ERR_INVALID, "too many coefficients");
  return PDEOPT_OK;
}

extern "C" pdeopt_status pdeopt_ch3d_rhs(const pdeopt_ch3d_desc* d, const float* u_dev, const float* halo_lo_dev,
                                         const float* halo_hi_dev, float* mu_work_dev, float* f_dev, int32_t batch,
                                         void* stream) {
  PdeoptDeviceGuard device_guard_(u_dev);
  pdeopt_status s = ch3d_check(d, batch);
  if (s != PDEOPT_OK) return s;
  if (!u_dev || !mu_work_dev || !f_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if ((halo_lo_dev || halo_hi_dev) && batch != 1) return fail(PDEOPT_ERR_INVALID, "slab halos need batch == 1");
  if ((halo_lo_dev == nullptr) != (halo_hi_dev == nullptr)) return fail(PDEOPT_ERR_INVALID, "give both halos or neither");
  Ch3dParams p;
  std::memset(&p, 0, sizeof(p));
  p.nx = d->nx; p.ny = d->ny; p.nz = d->nz; p.batch = batch;
  p.u = u_dev; p.halo_lo = halo_lo_dev; p.halo_hi = halo_hi_dev; p.mu = mu_work_dev; p.f = f_dev;
  p.inv_hx = (float)(1.0 / d->hx); p.inv_hy = (float)(1.0 / d->hy); p.inv_hz = (float)(1.0 / d->hz);
  p.inv_hx2 = (float)(1.0 / (d->hx * d->hx)); p.inv_hy2 = (float)(1.0 / (d->hy * d->hy)); p.inv_hz2 = (float)(1.0 / (d->hz * d->hz));
  p.kappa = (float)d->kappa;
  p.pw.mu_family = d->mu_family; p.pw.mu_ncoef = d->mu_ncoef; p.pw.mob_family = d->mob_family; p.pw.mob_ncoef = d->mob_ncoef;
  for (int i = 0; i < PDEOPT_MAX_COEF; ++i) { p.pw.mu_coef[i] = (float)d->mu_coef[i]; p.pw.mob_coef[i] = (float)d->mob_coef[i]; }
  const int bx = d->nz >= 256 ? 256 : (d->nz >= 128 ? 128 : (d->nz >= 64 ? 64 : 32));
  dim3 block(bx), g1((d->nz + bx - 1) / bx, d->ny, batch * (d->nx + 2)), g2((d->nz + bx - 1) / bx, d->ny, batch * d->nx);
  cudaStream_t st = (cudaStream_t)stream;
  // tile of the marching kernels: 16 x 64 (y, z), or 32 x 32 for small grids
  const bool tile64 = d->ny % kC3TY == 0 && d->nz % kC3TZ == 0;
  const bool tile32 = !tile64 && d->ny % 32 == 0 && d->nz % 32 == 0;
  const int tyy = tile64 ? kC3TY : 32, tzz = tile64 ? kC3TZ : 32;
  // planes marched per CTA: the longest chunk that still fills the GPU (3 CTAs per SM); short chunks
  // re-read 3 planes per chunk but a 64^3 or 128^3 domain would otherwise run on 4 or 32 CTAs
  int xl = 0;
  if (tile64 || tile32) {
    const int64_t tiles = (int64_t)(d->ny / tyy) * (d->nz / tzz) * batch;
    for (int c : {64, 32, 16, 8}) {
      if (d->nx % c != 0) continue;
      xl = c;
      if (tiles * (d->nx / c) >= 3 * 148) break;
    }
    // experiment hook: PDEOPT_CH3D_XL forces the chunk length (must divide nx)
    static const int force_xl = [] { const char* e = std::getenv("PDEOPT_CH3D_XL"); return e ? std::atoi(e) : 0; }();
    if (force_xl > 0 && d->nx % force_xl == 0) xl = force_xl;
  }
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool aligned = aligned16(u_dev) && aligned16(f_dev) && aligned16(halo_lo_dev) && aligned16(halo_hi_dev);
  if (xl > 0 && (int64_t)batch * (d->nx / xl) <= 65535 && (aligned || tile64)) {
    // fused 2.5-D marching kernel: one read of u, one write of f
    dim3 grid(d->nz / tzz, d->ny / tyy, batch * (d->nx / xl));
    if (aligned) {
      // register-marching version (128-bit accesses), specialised for the closure families with packed forms
      const int mf = d->mu_family, bf = d->mob_family;
#define PDEOPT_MARCH(MU_, MOB_)                                                                          \
  do {                                                                                                   \
    if (tile64) ch3d_rhs_march_kernel<MU_, MOB_, kC3TY, kC3TZ><<<grid, kC3Threads, 0, st>>>(p, xl);      \
    else ch3d_rhs_march_kernel<MU_, MOB_, 32, 32><<<grid, kC3Threads, 0, st>>>(p, xl);                   \
  } while (0)
      if (mf == MU_LOG && bf == MOB_CONST) PDEOPT_MARCH(MU_LOG, MOB_CONST);
      else if (mf == MU_LOG && bf == MOB_DEGENERATE) PDEOPT_MARCH(MU_LOG, MOB_DEGENERATE);
      else if (mf == MU_DOUBLE_WELL && bf == MOB_CONST) PDEOPT_MARCH(MU_DOUBLE_WELL, MOB_CONST);
      else if (bf == MOB_CONST) PDEOPT_MARCH(MU_RUNTIME, MOB_CONST);
      else PDEOPT_MARCH(MU_RUNTIME, MOB_RUNTIME);
#undef PDEOPT_MARCH
    } else {
      ch3d_rhs_fused_kernel<<<grid, kC3Threads, 0, st>>>(p, xl);
    }
    g_launches.fetch_add(1);
  } else {
    ch3d_mu_kernel<<<g1, block, 0, st>>>(p);
    ch3d_div_kernel<<<g2, block, 0, st>>>(p);
    g_launches.fetch_add(2);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("ch3d_rhs: ") + cudaGetErrorString(e));
  return PDEOPT_OK;
}

extern "C" int64_t pdeopt_ch3d_work_floats(const pdeopt_ch3d_desc* d, int32_t batch) {
  if (!d || batch <= 0) return 0;
  const int64_t pl = (int64_t)d->ny * d->nz;
  return (int64_t)batch * ((d->nx + 2) * pl + d->nx * pl + 2 * (int64_t)d->nx * d->ny * (d->nz / 2 + 1));
}

extern "C" pdeopt_status pdeopt_ch3d_step(const pdeopt_ch3d_desc* d, const float* y0_dev, float* y1_dev, int32_t batch,
                                          int32_t ksteps, const float* dt_host, const float* symbol_pos_dev,
                                          float* work_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(y0_dev);
  pdeopt_status s = ch3d_check(d, batch);
  if (s != PDEOPT_OK) return s;
  if (!y0_dev || !y1_dev || !dt_host || !symbol_pos_dev || !work_dev) return fail(PDEOPT_ERR_INVALID, "null argument");
  if (ksteps <= 0) return fail(PDEOPT_ERR_INVALID, "ksteps must be positive");
  if (!lf_size_ok(d->nx) || !lf_size_ok(d->ny) || !lf_size_ok(d->nz))
    return fail(PDEOPT_ERR_UNSUPPORTED, "ch3d_step: nx, ny, nz must be powers of two in [8, 512]");
  const int64_t nx = d->nx, ny = d->ny, nz = d->nz, pl = ny * nz, vol = nx * pl;
  const int64_t hp = nz / 2 + 1, hpl = ny * hp, hvol = nx * hpl;  // half spectrum along z (real field)
  float* mu = work_dev;
  float* f = mu + (int64_t)batch * (nx + 2) * pl;
  float* W = f + (int64_t)batch * vol;  // complex [batch][nx][ny][nz/2+1]
  pdeopt_line_geom gy{(int64_t)batch * nx * hp, hp, hpl, 1, (int32_t)ny, 0, 0, hp};
  pdeopt_line_geom gx{(int64_t)batch * hpl, hpl, hvol, 1, (int32_t)nx, 0, 0, hpl};
  pdeopt_line_geom gsym = gx;
  gsym.outer = 0;  // the symbol is shared by the batch
  const float* src = y0_dev;
  for (int k = 0; k < ksteps; ++k) {
    if ((s = pdeopt_ch3d_rhs(d, src, nullptr, nullptr, mu, f, batch, stream)) != PDEOPT_OK) return s;
    if ((s = pdeopt_fft_lines_r2c(f, W, (int32_t)nz, (int64_t)batch * nx * ny, stream)) != PDEOPT_OK) return s;
    if ((s = pdeopt_fft_lines(W, W, (int32_t)ny, &gy, &gy, 0, 0, 1.0f, stream)) != PDEOPT_OK) return s;
    if ((s = pdeopt_fft_lines_imex(W, W, (int32_t)nx, &gx, symbol_pos_dev, &gsym, dt_host[k], 1.0f / (float)vol, stream)) != PDEOPT_OK) return s;
    if ((s = pdeopt_fft_lines(W, W, (int32_t)ny, &gy, &gy, 1, 0, 1.0f, stream)) != PDEOPT_OK) return s;
    if ((s = pdeopt_fft_lines_c2r_update(W, (int32_t)nz, (int64_t)batch * nx * ny, src, y1_dev, dt_host[k], stream)) != PDEOPT_OK) return s;
    src = y1_dev;
  }
  return PDEOPT_OK;
}

extern "C" int64_t pdeopt_ch3d_adjoint_work_floats(const pdeopt_ch3d_desc* d, int32_t batch) {
  if (!d || batch <= 0) return 0;
  const int64_t vol = (int64_t)d->nx * d->ny * d->nz;
  return (int64_t)batch * (6 * vol + 2 * (int64_t)d->nx * d->ny * (d->nz / 2 + 1));
}

extern "C" pdeopt_status pdeopt_ch3d_adjoint_step(const pdeopt_ch3d_desc* d, const float* u_dev, const float* lam1_dev,
                                                  float* lam0_dev, int32_t batch, float dt, const float* symbol_pos_dev,
                                                  float* work_dev, double* gmu_dev, double* gmob_dev, void* stream) {
  PdeoptDeviceGuard device_guard_(u_dev);
  pdeopt_status s = ch3d_check(d, batch);
  if (s != PDEOPT_OK) return s;
  if (!u_dev || !lam1_dev || !lam0_dev || !symbol_pos_dev || !work_dev || !gmu_dev || !gmob_dev)
    return fail(PDEOPT_ERR_INVALID, "null argument");
  if (!lf_size_ok(d->nx) || !lf_size_ok(d->ny) || !lf_size_ok(d->nz))
    return fail(PDEOPT_ERR_UNSUPPORTED, "ch3d adjoint: nx, ny, nz must be powers of two in [8, 512]");
  const int64_t nx = d->nx, ny = d->ny, nz = d->nz, vol = nx * ny * nz, n = vol * batch;
  const int64_t hp = nz / 2 + 1, hpl = ny * hp, hvol = nx * hpl;
  float* zeros = work_dev;
  float* w = work_dev + n;
  float* W = work_dev + 6 * n;  // complex half spectrum
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(zeros, 0, sizeof(float) * n, st));
  // w = dt G lam1: the forward step's spectral filter applied to lam1 with a zero state
  pdeopt_line_geom gy{(int64_t)batch * nx * hp, hp, hpl, 1, (int32_t)ny, 0, 0, hp};
  pdeopt_line_geom gx{(int64_t)batch * hpl, hpl, hvol, 1, (int32_t)nx, 0, 0, hpl};
  pdeopt_line_geom gsym = gx;
  gsym.outer = 0;
  if ((s = pdeopt_fft_lines_r2c(lam1_dev, W, (int32_t)nz, (int64_t)batch * nx * ny, stream)) != PDEOPT_OK) return s;
  if ((s = pdeopt_fft_lines(W, W, (int32_t)ny, &gy, &gy, 0, 0, 1.0f, stream)) != PDEOPT_OK) return s;
  if ((s = pdeopt_fft_lines_imex(W, W, (int32_t)nx, &gx, symbol_pos_dev, &gsym, dt, 1.0f / (float)vol, stream)) != PDEOPT_OK) return s;
  if ((s = pdeopt_fft_lines(W, W, (int32_t)ny, &gy, &gy, 1, 0, 1.0f, stream)) != PDEOPT_OK) return s;
  if ((s = pdeopt_fft_lines_c2r_update(W, (int32_t)nz, (int64_t)batch * nx * ny, zeros, w, dt, stream)) != PDEOPT_OK) return s;
  Ch3AdjParams p;
  std::memset(&p, 0, sizeof(p));
  p.nx = d->nx; p.ny = d->ny; p.nz = d->nz; p.batch = batch;
  p.u = u_dev; p.w = w; p.lam1 = lam1_dev; p.lam0 = lam0_dev;
  p.mu = work_dev + 2 * n; p.dd = work_dev + 3 * n; p.mub = work_dev + 4 * n; p.db = work_dev + 5 * n;
  p.gmu = gmu_dev; p.gmob = gmob_dev;
  p.inv_hx = (float)(1.0 / d->hx); p.inv_hy = (float)(1.0 / d->hy); p.inv_hz = (float)(1.0 / d->hz);
  p.inv_hx2 = (float)(1.0 / (d->hx * d->hx)); p.inv_hy2 = (float)(1.0 / (d->hy * d->hy)); p.inv_hz2 = (float)(1.0 / (d->hz * d->hz));
  p.kappa = (float)d->kappa;
  p.pw.mu_family = d->mu_family; p.pw.mu_ncoef = d->mu_ncoef; p.pw.mob_family = d->mob_family; p.pw.mob_ncoef = d->mob_ncoef;
  for (int i = 0; i < PDEOPT_MAX_COEF; ++i) { p.pw.mu_coef[i] = (float)d->mu_coef[i]; p.pw.mob_coef[i] = (float)d->mob_coef[i]; }
  dim3 grid((unsigned)((vol + 255) / 256), batch);
  ch3_adj_mu_kernel<<<grid, 256, 0, st>>>(p);
  ch3_adj_bar_kernel<<<grid, 256, 0, st>>>(p);
  ch3_adj_out_kernel<<<grid, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PDEOPT_ERR_CUDA, std::string("ch3d adjoint step: ") + cudaGetErrorString(e));
  g_launches.fetch_add(3);
  return PDEOPT_OK;
}

