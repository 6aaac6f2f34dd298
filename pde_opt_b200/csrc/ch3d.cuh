// 3-D Cahn-Hilliard on a periodic grid: finite-difference RHS kernels and the line-FFT functors of
// the semi-implicit step (sm_100a).  HBM-bound streaming kernels: the 512^3 field (512 MB) does not
// fit on chip, so the design goal is the fewest passes over HBM, all of them coalesced.
//
// Replaces CahnHilliard3DPeriodic.rhs_fd (pde_opt/numerics/equations/cahn_hilliard.py:177-200) with
// the 3-D stencils of pde_opt/numerics/utils/derivatives.py:15-21, :29-36, :44-51, :59-66, and
// SemiImplicitFourierSpectral.step (solvers.py:56-70) on 3-D fields.
//
// Slab decomposition (BASELINE config 5): a rank holds nx planes [nx][ny][nz] of the global field;
// the two planes on either side come in through halo arrays (null = periodic wrap on this rank).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "linefft.cuh"
#include "pointwise.cuh"

namespace pdeopt {

struct Ch3dParams {
  int nx, ny, nz, batch;
  const float* u;        // [batch][nx][ny][nz]
  const float* halo_lo;  // [2][ny][nz]: planes x = -2, -1 (or null)
  const float* halo_hi;  // [2][ny][nz]: planes x = nx, nx+1 (or null)
  float* mu;             // [nx+2][ny][nz]: planes x = -1 .. nx
  float* f;              // [nx][ny][nz]
  float inv_hx, inv_hy, inv_hz, inv_hx2, inv_hy2, inv_hz2, kappa;
  PointwiseParams pw;
};

// plane x of domain b; halos (slab mode) are only meaningful for a single domain
__device__ __forceinline__ const float* ch3d_plane(const Ch3dParams& p, int b, int x) {
  const size_t pl = (size_t)p.ny * p.nz;
  const float* u = p.u + (size_t)b * p.nx * pl;
  if (x < 0) return p.halo_lo ? p.halo_lo + (size_t)(x + 2) * pl : u + (size_t)(x + p.nx) * pl;
  if (x >= p.nx) return p.halo_hi ? p.halo_hi + (size_t)(x - p.nx) * pl : u + (size_t)(x - p.nx) * pl;
  return u + (size_t)x * pl;
}

// mu = mu_h(u) - kappa * lap(u) on planes x = -1 .. nx   (cahn_hilliard.py:179; derivatives.py:15-21)
__global__ void __launch_bounds__(256) ch3d_mu_kernel(const __grid_constant__ Ch3dParams p) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z / (p.nx + 2);
  const int xm = blockIdx.z % (p.nx + 2);  // 0 .. nx+1  <->  x = xm - 1
  if (z >= p.nz) return;
  const int x = xm - 1;
  const float* c0 = ch3d_plane(p, b, x);
  const float* cm = ch3d_plane(p, b, x - 1);
  const float* cp = ch3d_plane(p, b, x + 1);
  const int yp = (y + 1 == p.ny) ? 0 : y + 1, ym = (y == 0) ? p.ny - 1 : y - 1;
  const int zp = (z + 1 == p.nz) ? 0 : z + 1, zm = (z == 0) ? p.nz - 1 : z - 1;
  const size_t o = (size_t)y * p.nz + z;
  const float u = c0[o];
  const float lap = ((cp[o] - 2.0f * u) + cm[o]) * p.inv_hx2 +
                    ((c0[(size_t)yp * p.nz + z] - 2.0f * u) + c0[(size_t)ym * p.nz + z]) * p.inv_hy2 +
                    ((c0[(size_t)y * p.nz + zp] - 2.0f * u) + c0[(size_t)y * p.nz + zm]) * p.inv_hz2;
  p.mu[((size_t)b * (p.nx + 2) + xm) * p.ny * p.nz + o] = mu_h<MU_RUNTIME>(u, p.pw, 0.0f) - p.kappa * lap;
}

// f = div( avg_face(D(u)) * grad_face(mu) )   (cahn_hilliard.py:181-200)
__global__ void __launch_bounds__(256) ch3d_div_kernel(const __grid_constant__ Ch3dParams p) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z / p.nx, x = blockIdx.z % p.nx;
  if (z >= p.nz) return;
  const size_t pl = (size_t)p.ny * p.nz;
  const float* u0 = ch3d_plane(p, b, x);
  const float* um = ch3d_plane(p, b, x - 1);
  const float* up = ch3d_plane(p, b, x + 1);
  const float* m0 = p.mu + ((size_t)b * (p.nx + 2) + x + 1) * pl;
  const float* mm = m0 - pl;
  const float* mp = m0 + pl;
  const int yp = (y + 1 == p.ny) ? 0 : y + 1, ym = (y == 0) ? p.ny - 1 : y - 1;
  const int zp = (z + 1 == p.nz) ? 0 : z + 1, zm = (z == 0) ? p.nz - 1 : z - 1;
  const size_t o = (size_t)y * p.nz + z;
  const size_t oyp = (size_t)yp * p.nz + z, oym = (size_t)ym * p.nz + z;
  const size_t ozp = (size_t)y * p.nz + zp, ozm = (size_t)y * p.nz + zm;
  const float D0 = mob<MOB_RUNTIME>(u0[o], p.pw), mu0 = m0[o];
  auto flux = [&](float Dn, float mun, float inv_h) { return (0.5f * (D0 + Dn)) * ((mun - mu0) * inv_h); };
  auto fluxb = [&](float Dn, float mun, float inv_h) { return (0.5f * (Dn + D0)) * ((mu0 - mun) * inv_h); };
  const float fx = (flux(mob<MOB_RUNTIME>(up[o], p.pw), mp[o], p.inv_hx) - fluxb(mob<MOB_RUNTIME>(um[o], p.pw), mm[o], p.inv_hx)) * p.inv_hx;
  const float fy = (flux(mob<MOB_RUNTIME>(u0[oyp], p.pw), m0[oyp], p.inv_hy) - fluxb(mob<MOB_RUNTIME>(u0[oym], p.pw), m0[oym], p.inv_hy)) * p.inv_hy;
  const float fz = (flux(mob<MOB_RUNTIME>(u0[ozp], p.pw), m0[ozp], p.inv_hz) - fluxb(mob<MOB_RUNTIME>(u0[ozm], p.pw), m0[ozm], p.inv_hz)) * p.inv_hz;
  p.f[((size_t)b * p.nx + x) * pl + o] = (fx + fy) + fz;
}

// ---- fused single-pass RHS: 2.5-D marching ----------------------------------------------------
// One CTA owns a (y, z) tile of 16 x 64 points and marches along x for `xl` planes, so u is read
// from HBM once (plus the tile halo, which neighbouring CTAs share through L2) and f written once:
// ~2 field passes instead of the 5 of the two-kernel version above.  Rolling shared-memory windows:
// three u planes (tile + halo 2) and two (mu, D) planes (tile + ring 1); the x-face flux of the
// previous plane is carried in registers.
constexpr int kC3TY = 16, kC3TZ = 64, kC3Threads = 256;
constexpr int kC3UY = kC3TY + 4, kC3UZ = kC3TZ + 4, kC3UZP = 72;
constexpr int kC3MY = kC3TY + 2, kC3MZ = kC3TZ + 2, kC3MZP = 68;

struct Ch3dSmem {
  float U[3][kC3UY][kC3UZP];
  float2 MD[2][kC3MY][kC3MZP];
};

__global__ void __launch_bounds__(kC3Threads) ch3d_rhs_fused_kernel(const __grid_constant__ Ch3dParams p, int xl) {
  __shared__ Ch3dSmem S;
  const int tid = threadIdx.x;
  const int z0 = blockIdx.x * kC3TZ, y0 = blockIdx.y * kC3TY;
  const int nchunk = p.nx / xl;
  const int b = blockIdx.z / nchunk, x0 = (blockIdx.z % nchunk) * xl;
  // this thread's 4 points of row ty: z = tz + 16 j, so that the 16 lanes of a half-warp read
  // consecutive (mu, D) pairs (stride-4 ownership made every shared-memory access 4-way conflicted)
  const int ty = tid >> 4, tz = tid & 15;
  constexpr int kULoads = (kC3UY * kC3UZ + kC3Threads - 1) / kC3Threads;  // 6

  // global offsets (within a plane) of the halo-2 tile elements this thread stages
  int goff[kULoads], soff[kULoads];
#pragma unroll
  for (int i = 0; i < kULoads; ++i) {
    const int e = tid + i * kC3Threads;
    if (e < kC3UY * kC3UZ) {
      const int yy = e / kC3UZ, zz = e - yy * kC3UZ;
      int gy = y0 + yy - 2, gz = z0 + zz - 2;
      gy = gy < 0 ? gy + p.ny : (gy >= p.ny ? gy - p.ny : gy);
      gz = gz < 0 ? gz + p.nz : (gz >= p.nz ? gz - p.nz : gz);
      goff[i] = gy * p.nz + gz;
      soff[i] = yy * kC3UZP + zz;
    } else {
      goff[i] = -1;
      soff[i] = 0;
    }
  }
  auto fetch = [&](int x, float (&r)[kULoads]) {
    const float* pl = ch3d_plane(p, b, x);
#pragma unroll
    for (int i = 0; i < kULoads; ++i) r[i] = goff[i] >= 0 ? __ldg(pl + goff[i]) : 0.f;
  };
  auto stage = [&](int slot, const float (&r)[kULoads]) {
    float* u = &S.U[slot][0][0];
#pragma unroll
    for (int i = 0; i < kULoads; ++i)
      if (goff[i] >= 0) u[soff[i]] = r[i];
  };
  // (mu, D) of plane c on tile + ring 1 from U slots (c-1, c, c+1) = (sm, s0, sp)
  auto mu_plane = [&](int sm, int s0, int sp, int md) {
    for (int e = tid; e < kC3MY * kC3MZ; e += kC3Threads) {
      const int yy = e / kC3MZ, zz = e - yy * kC3MZ;  // ring coordinates; U index = +1
      const float u = S.U[s0][yy + 1][zz + 1];
      const float lap = ((S.U[sp][yy + 1][zz + 1] - 2.0f * u) + S.U[sm][yy + 1][zz + 1]) * p.inv_hx2 +
                        ((S.U[s0][yy + 2][zz + 1] - 2.0f * u) + S.U[s0][yy][zz + 1]) * p.inv_hy2 +
                        ((S.U[s0][yy + 1][zz + 2] - 2.0f * u) + S.U[s0][yy + 1][zz]) * p.inv_hz2;
      S.MD[md][yy][zz] = make_float2(mu_h<MU_RUNTIME>(u, p.pw, 0.0f) - p.kappa * lap, mob<MOB_RUNTIME>(u, p.pw));
    }
  };

  float r[kULoads];
  // prologue: planes x0-2, x0-1 -> slots 0, 1; plane x0 prefetched
  fetch(x0 - 2, r);
  stage(0, r);
  fetch(x0 - 1, r);
  stage(1, r);
  fetch(x0, r);
  float fx_prev[4] = {0.f, 0.f, 0.f, 0.f};
  // iteration it: current plane c = x0 + it; U slots: (c-1) -> (it+1)%3 ... kept as rolling indices
  int sm = 0, s0 = 1, sp = 2;  // slots of planes c-1, c, c+1 at it = -1 (c = x0 - 1)
  int md_prev = 0, md_cur = 1;
  for (int it = -1; it <= xl; ++it) {
    const int c = x0 + it;
    stage(sp, r);                      // plane c + 1
    if (it < xl) fetch(c + 2, r);      // prefetch plane c + 2 while this iteration computes
    __syncthreads();
    mu_plane(sm, s0, sp, md_cur);
    __syncthreads();
    if (it >= 0) {
      // x-face flux between planes c-1 and c, own points (cahn_hilliard.py:181-186)
      float fx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = S.MD[md_prev][ty + 1][tz + 1 + 16 * j], bb = S.MD[md_cur][ty + 1][tz + 1 + 16 * j];
        fx[j] = (0.5f * (a.y + bb.y)) * ((bb.x - a.x) * p.inv_hx);
      }
      if (it >= 1) {
        // f of plane c-1: x-divergence + in-plane flux divergence from the (mu, D) plane with ring
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int yy = ty + 1, zz = tz + 1 + 16 * j;
          const float2 m0 = S.MD[md_prev][yy][zz];
          const float2 myp = S.MD[md_prev][yy + 1][zz], mym = S.MD[md_prev][yy - 1][zz];
          const float2 mzp = S.MD[md_prev][yy][zz + 1], mzm = S.MD[md_prev][yy][zz - 1];
          const float fyp = (0.5f * (m0.y + myp.y)) * ((myp.x - m0.x) * p.inv_hy);
          const float fym = (0.5f * (mym.y + m0.y)) * ((m0.x - mym.x) * p.inv_hy);
          const float fzp = (0.5f * (m0.y + mzp.y)) * ((mzp.x - m0.x) * p.inv_hz);
          const float fzm = (0.5f * (mzm.y + m0.y)) * ((m0.x - mzm.x) * p.inv_hz);
          out[j] = ((fx[j] - fx_prev[j]) * p.inv_hx + (fyp - fym) * p.inv_hy) + (fzp - fzm) * p.inv_hz;
        }
        float* dst = p.f + ((size_t)b * p.nx + (c - 1)) * p.ny * p.nz + (size_t)(y0 + ty) * p.nz + z0 + tz;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[16 * j] = out[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) fx_prev[j] = fx[j];
    }
    __syncthreads();  // all reads of U slot `sm` and MD[md_prev] done before they are overwritten
    const int t = sm;
    sm = s0;
    s0 = sp;
    sp = t;
    md_prev ^= 1;
    md_cur ^= 1;
  }
}

// ---- register-marching single-pass RHS ----------------------------------------------------------
// Same tile (16 x 64 points of a (y, z) plane per 256-thread CTA, marching along x) as the kernel
// above, but every thread keeps the x-column of its four consecutive z points in registers and
// shared memory only carries what NEIGHBOURS need: planes of u (tile + halo 2) and of (mu, D)
// (tile + ring 1).  u planes arrive by cp.async into a four-stage ring (planes c .. c+3: three planes
// per CTA in flight, which is what it takes to cover HBM latency at three CTAs per SM); (mu, D) are
// double buffered, so there is one barrier per plane.  128-bit shared and global accesses,
// z-neighbours inside a warp by shuffle, packed f32x2 arithmetic; 20 lanes of every warp additionally
// march one ring point each.  The loop body is instantiated per ring stage so that every
// shared-memory access is base register + immediate.
//   f = cx (Gx - Gx_prev) + cy (Gy+ - Gy-) + cz (Gz+ - Gz-),  G = (D + D_nbr) (mu_nbr - mu), c = 1/(2h^2)
// (cahn_hilliard.py:177-200 with the constant factors collected as in sifs128.cuh; a few ulp).
// Tile shapes: 16 x 64 (y, z) for nz % 64 == 0, 32 x 32 for the small grids (nz % 32 == 0, ny % 32 == 0).
constexpr int kM3NS = 4;  // u stages
template <int TY, int TZ>
struct Ch3dMarchSmem {
  // row pitch in floats: 4 halo columns | TZ | halo; 76 mod 32 = 12 keeps the column-ring accesses (one per
  // row) at two per bank
  static constexpr int P = TZ == 64 ? 76 : TZ + 8;  // (the 32 x 32 tile would pass 48 KB of static shared memory with +12)
  static constexpr int UR = TY + 4;  // u rows (halo 2)
  static constexpr int MR = TY + 2;  // mu / D rows (ring 1)
  float U[kM3NS][UR][P];  // [y_local + 2][z_local + 4]
  float Mu[2][MR][P];     // [y_local + 1][z_local + 4]
  float Dm[2][MR][P];
};

__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 ld4(const float* q) { return *reinterpret_cast<const float4*>(q); }
__device__ __forceinline__ void cp_async16(float* smem, const float* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem, const float* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MU, int MOB, int TY, int TZ>
__global__ void __launch_bounds__(kC3Threads, 3) ch3d_rhs_march_kernel(const __grid_constant__ Ch3dParams p, int xl) {
  using Smem = Ch3dMarchSmem<TY, TZ>;
  __shared__ __align__(16) Smem S;
  constexpr int kM3P = Smem::P, kM3US = Smem::UR * Smem::P, kM3MS = Smem::MR * Smem::P;
  constexpr int LZ = TZ / 4;  // lanes per tile row (four consecutive z points per thread)
  static_assert(TY * LZ == kC3Threads, "tile must give every thread four points");
  constexpr int kR1 = 2 * TZ + 2 * TY, kR2 = kR1 + 4;  // ring-1 points; ring-2 points + corners
  constexpr int kR1W = kR1 / 8, kR2W = (kR2 + 7) / 8;  // per warp
  constexpr bool DCONST = (MOB == MOB_CONST);
  const int tid = threadIdx.x;
  const int z0 = blockIdx.x * TZ, y0 = blockIdx.y * TY;
  const int nchunk = p.nx / xl;
  const int b = blockIdx.z / nchunk, x0 = (blockIdx.z % nchunk) * xl;
  const int ty = tid / LZ, tq = tid % LZ;
  const int own_off = (y0 + ty) * p.nz + z0 + 4 * tq;
  const size_t pl = (size_t)p.ny * p.nz;

  // ring roles, spread evenly over the 8 warps so that no warp is late at the barrier: the first lanes of
  // every warp (20 for the 16 x 64 tile) march one ring-1 point (mu, D needed there), the first 21 stage
  // one ring-2 / corner point of u (only neighbours read it)
  auto wrap_off = [&](int yl, int zl) {
    int gy = y0 + yl, gz = z0 + zl;
    gy = gy < 0 ? gy + p.ny : (gy >= p.ny ? gy - p.ny : gy);
    gz = gz < 0 ? gz + p.nz : (gz >= p.nz ? gz - p.nz : gz);
    return gy * p.nz + gz;
  };
  const int lane = tid & 31, warp = tid >> 5;
  const bool r1 = lane < kR1W;
  const int ri = warp * kR1W + lane;
  int r1y = 0, r1z = 0;
  if (ri < TZ) { r1y = -1; r1z = ri; }
  else if (ri < 2 * TZ) { r1y = TY; r1z = ri - TZ; }
  else if (ri < 2 * TZ + TY) { r1y = ri - 2 * TZ; r1z = -1; }
  else if (ri < kR1) { r1y = ri - 2 * TZ - TY; r1z = TZ; }
  const int r1_off = wrap_off(r1y, r1z);
  const int rj = warp * kR2W + lane;
  const bool r2 = lane < kR2W && rj < kR2;
  int r2y = 0, r2z = 0;
  if (r2) {
    const int j = rj;
    if (j < TZ) { r2y = -2; r2z = j; }
    else if (j < 2 * TZ) { r2y = TY + 1; r2z = j - TZ; }
    else if (j < 2 * TZ + TY) { r2y = j - 2 * TZ; r2z = -2; }
    else if (j < kR1) { r2y = j - 2 * TZ - TY; r2z = TZ + 1; }
    else { r2y = ((j - kR1) & 2) ? TY : -1; r2z = ((j - kR1) & 1) ? TZ : -1; }
  }
  const int r2_off = r2 ? wrap_off(r2y, r2z) : 0;
  // per-thread shared-memory bases (stage / buffer 0); the others are compile-time offsets away
  float* const Uown = &S.U[0][ty + 2][4 + 4 * tq];
  float* const Ur1 = &S.U[0][r1y + 2][r1z + 4];
  float* const Ur2 = &S.U[0][r2y + 2][r2z + 4];
  float* const Mown = &S.Mu[0][ty + 1][4 + 4 * tq];
  float* const Mr1 = &S.Mu[0][r1y + 1][r1z + 4];
  constexpr int kMD = 2 * kM3MS;  // Mu -> Dm

  const float2 ihx2 = splat2(p.inv_hx2), ihy2 = splat2(p.inv_hy2), ihz2 = splat2(p.inv_hz2), m2 = splat2(-2.0f);
  const float2 mkappa = splat2(-p.kappa), zero2 = make_float2(0.f, 0.f);
  const float dscale = DCONST ? 2.0f * p.pw.mob_coef[0] : 1.0f;  // (D + D) folded into the face coefficient
  const float2 cx = splat2(0.5f * p.inv_hx * p.inv_hx * dscale), cy = splat2(0.5f * p.inv_hy * p.inv_hy * dscale),
               cz = splat2(0.5f * p.inv_hz * p.inv_hz * dscale);

  // mu = mu_h(u) - kappa lap(u) (derivatives.py:15-21, cahn_hilliard.py:179) and D(u) for a packed pair
  auto mu_of = [&](float2 uc, float2 uxp, float2 uxm, float2 uyp, float2 uym, float2 uzp, float2 uzm, float2& mu, float2& D) {
    const float2 dxx = add2(fma2(uc, m2, uxp), uxm);
    const float2 dyy = add2(fma2(uc, m2, uyp), uym);
    const float2 dzz = add2(fma2(uc, m2, uzp), uzm);
    const float2 lap = fma2(dzz, ihz2, fma2(dyy, ihy2, mul2(dxx, ihx2)));
    float2 mh;
    mu_mob_pair<MU, MOB>(uc, p.pw, zero2, mh, D);
    mu = fma2(lap, mkappa, mh);
  };

  // asynchronous plane copies: the plane pointer advances by one plane per issue and is re-derived
  // only where the periodic wrap / slab halo changes the array.  One commit group per plane (empty
  // past the last plane, so the wait counts stay uniform).
  int xnext = x0 - 2;
  const int xlast = x0 + xl + 1;
  const float* pnext = ch3d_plane(p, b, xnext);
  auto issue = [&](int stage_off) {
    if (xnext <= xlast) {
      cp_async16(Uown + stage_off, pnext + own_off);
      if (r1) cp_async4(Ur1 + stage_off, pnext + r1_off);
      if (r2) cp_async4(Ur2 + stage_off, pnext + r2_off);
      ++xnext;
      pnext += pl;
      if (xnext == 0 || xnext == p.nx) pnext = ch3d_plane(p, b, xnext);
    }
    cp_async_commit();
  };
  issue(0 * kM3US);  // x0 - 2
  issue(1 * kM3US);  // x0 - 1
  issue(2 * kM3US);  // x0
  issue(3 * kM3US);  // x0 + 1
  cp_async_wait<2>();
  // column registers: planes c-1, c (c+1 is read from its stage when needed)
  float4 um = ld4(Uown), u0 = ld4(Uown + kM3US), up;
  float rum = 0.f, ru0 = 0.f, rup = 0.f;
  if (r1) {
    rum = Ur1[0];
    ru0 = Ur1[kM3US];
  }
  __syncthreads();     // plane x0 - 1 visible to everybody; stage 0 free again
  issue(0 * kM3US);    // x0 + 2

  float2 mu_p[2] = {zero2, zero2}, D_p[2] = {zero2, zero2}, gx_p[2] = {zero2, zero2}, dy_p[2] = {zero2, zero2},
         dz_p[2] = {zero2, zero2};
  float* fout = p.f + ((size_t)b * p.nx + x0) * pl + own_off;  // plane x0 is emitted at it = 1
  const bool edgeL = tq == 0, edgeR = tq == LZ - 1;

  // one plane c held in stage ST (planes c+1 .. c+3 follow in the ring)
  auto plane = [&](auto st_c, int it) {
    constexpr int ST = decltype(st_c)::value;
    constexpr int PU = ST * kM3US, PUn = ((ST + 1) % kM3NS) * kM3US;
    constexpr int PM = (ST & 1) * kM3MS;
    cp_async_wait<2>();  // this thread's part of plane c+1 has landed
    up = ld4(Uown + PUn);
    if (r1) rup = Ur1[PUn];
    // ---- mu, D of plane c: own four points (packed pairs), then the ring point ----
    float2 mu[2], D[2];
    {
      const float4 uyp = ld4(Uown + PU + kM3P), uym = ld4(Uown + PU - kM3P);
      float uL = __shfl_up_sync(0xffffffffu, u0.w, 1), uR = __shfl_down_sync(0xffffffffu, u0.x, 1);
      if (edgeL) uL = Uown[PU - 1];
      if (edgeR) uR = Uown[PU + 4];
      const float2 uc[2] = {lo2(u0), hi2(u0)};
      const float2 uxp[2] = {lo2(up), hi2(up)}, uxm[2] = {lo2(um), hi2(um)};
      const float2 uyp2[2] = {lo2(uyp), hi2(uyp)}, uym2[2] = {lo2(uym), hi2(uym)};
      const float2 mid = make_float2(u0.y, u0.z);
      const float2 uzm[2] = {make_float2(uL, u0.x), mid}, uzp[2] = {mid, make_float2(u0.w, uR)};
#pragma unroll
      for (int h = 0; h < 2; ++h) mu_of(uc[h], uxp[h], uxm[h], uyp2[h], uym2[h], uzp[h], uzm[h], mu[h], D[h]);
      *reinterpret_cast<float4*>(Mown + PM) = make_float4(mu[0].x, mu[0].y, mu[1].x, mu[1].y);
      if constexpr (!DCONST) *reinterpret_cast<float4*>(Mown + PM + kMD) = make_float4(D[0].x, D[0].y, D[1].x, D[1].y);
    }
    if (r1) {
      // same packed expression as the owner of this point in the neighbouring tile: fluxes across tile
      // edges are bit-identical on both sides
      float2 rm, rD;
      mu_of(splat2(ru0), splat2(rup), splat2(rum), splat2(Ur1[PU + kM3P]), splat2(Ur1[PU - kM3P]), splat2(Ur1[PU + 1]),
            splat2(Ur1[PU - 1]), rm, rD);
      Mr1[PM] = rm.x;
      if constexpr (!DCONST) Mr1[PM + kMD] = rD.x;
    }
    // z-neighbours of mu inside the warp (tile edges come from the ring after the barrier)
    float mL = __shfl_up_sync(0xffffffffu, mu[1].y, 1), mR = __shfl_down_sync(0xffffffffu, mu[0].x, 1);
    float dL = 0.f, dR = 0.f;
    if constexpr (!DCONST) {
      dL = __shfl_up_sync(0xffffffffu, D[1].y, 1);
      dR = __shfl_down_sync(0xffffffffu, D[0].x, 1);
    }
    __syncthreads();  // (mu, D) of plane c complete; plane c+1 landed for everybody; stage ST free
    issue(PU);        // plane c+4

    // ---- fluxes: x-face between c-1 and c, in-plane divergence of plane c; emit f of plane c-1 ----
    {
      const float4 myp = ld4(Mown + PM + kM3P), mym = ld4(Mown + PM - kM3P);
      if (edgeL) mL = Mown[PM - 1];
      if (edgeR) mR = Mown[PM + 4];
      const float2 myp2[2] = {lo2(myp), hi2(myp)}, mym2[2] = {lo2(mym), hi2(mym)};
      const float2 mmid = make_float2(mu[0].y, mu[1].x);
      const float2 mzm[2] = {make_float2(mL, mu[0].x), mmid}, mzp[2] = {mmid, make_float2(mu[1].y, mR)};
      float2 Dyp2[2], Dym2[2], Dzm[2], Dzp[2];
      if constexpr (!DCONST) {
        const float4 dyp = ld4(Mown + PM + kMD + kM3P), dym = ld4(Mown + PM + kMD - kM3P);
        if (edgeL) dL = Mown[PM + kMD - 1];
        if (edgeR) dR = Mown[PM + kMD + 4];
        Dyp2[0] = lo2(dyp); Dyp2[1] = hi2(dyp); Dym2[0] = lo2(dym); Dym2[1] = hi2(dym);
        const float2 dmid = make_float2(D[0].y, D[1].x);
        Dzm[0] = make_float2(dL, D[0].x); Dzm[1] = dmid; Dzp[0] = dmid; Dzp[1] = make_float2(D[1].y, dR);
      }
      float2 o2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float2 gx, dy, dz;
        if constexpr (DCONST) {
          gx = sub2(mu[h], mu_p[h]);
          dy = sub2(sub2(myp2[h], mu[h]), sub2(mu[h], mym2[h]));
          dz = sub2(sub2(mzp[h], mu[h]), sub2(mu[h], mzm[h]));
        } else {
          gx = mul2(add2(D_p[h], D[h]), sub2(mu[h], mu_p[h]));
          dy = sub2(mul2(add2(D[h], Dyp2[h]), sub2(myp2[h], mu[h])), mul2(add2(Dym2[h], D[h]), sub2(mu[h], mym2[h])));
          dz = sub2(mul2(add2(D[h], Dzp[h]), sub2(mzp[h], mu[h])), mul2(add2(Dzm[h], D[h]), sub2(mu[h], mzm[h])));
        }
        // f(c-1) = (cx (gx - gx_prev) + cy dy_prev) + cz dz_prev
        o2[h] = fma2(dz_p[h], cz, fma2(dy_p[h], cy, mul2(sub2(gx, gx_p[h]), cx)));
        gx_p[h] = gx;
        dy_p[h] = dy;
        dz_p[h] = dz;
        mu_p[h] = mu[h];
        if constexpr (!DCONST) D_p[h] = D[h];
      }
      if (it >= 1) {
        *reinterpret_cast<float4*>(fout) = make_float4(o2[0].x, o2[0].y, o2[1].x, o2[1].y);
        fout += pl;
      }
    }
    um = u0; u0 = up;
    rum = ru0; ru0 = rup;
  };
  // it = -1 .. xl: xl + 2 planes, starting in stage 1 (xl is a multiple of 8, so 4k + 2 planes)
  int it = -1;
  for (; it + 3 <= xl; it += 4) {
    plane(std::integral_constant<int, 1>{}, it);
    plane(std::integral_constant<int, 2>{}, it + 1);
    plane(std::integral_constant<int, 3>{}, it + 2);
    plane(std::integral_constant<int, 0>{}, it + 3);
  }
  plane(std::integral_constant<int, 1>{}, it);
  plane(std::integral_constant<int, 2>{}, it + 1);
  cp_async_wait<0>();
}

// ---- line-FFT functors of the 3-D semi-implicit step -------------------------------------------
struct LfLoadReal {  // real array -> complex with zero imaginary part
  const float* p;
  LineGeom g;
  __device__ __forceinline__ LineGeom gin() const { return g; }
  __device__ __forceinline__ float2 load(long long off, long long, int) const { return make_float2(p[off], 0.f); }
};
// multiplier scale / (1 + dt * symbol) with the (A-folded, position-ordered) symbol laid out like the data
struct LfMidImex {
  const float* sym;
  LineGeom g;
  float dt, scale;
  __device__ __forceinline__ LineGeom gaux() const { return g; }
  __device__ __forceinline__ float2 apply(float2 v, long long off_aux, long long, int) const {
    const float m = __fdividef(scale, fmaf(dt, sym[off_aux], 1.0f));
    return make_float2(v.x * m, v.y * m);
  }
  using Aux = float;
  __device__ __forceinline__ float fetch(long long off_aux, long long) const { return __ldg(sym + off_aux); }
  __device__ __forceinline__ float2 apply_aux(float2 v, float sg, long long, int) const {
    const float m = __fdividef(scale, fmaf(dt, sg, 1.0f));
    return make_float2(v.x * m, v.y * m);
  }
};
// y1 = y0 + dt * Re(g)   (solvers.py:63)
struct LfStoreUpdate {
  const float* y0;
  float* y1;
  LineGeom g;
  float dt;
  __device__ __forceinline__ LineGeom gout() const { return g; }
  __device__ __forceinline__ float2 pre(long long off) const { return make_float2(y0[off], 0.f); }
  __device__ __forceinline__ void store(long long off, long long, int, float2 v, float2 pre) const { y1[off] = fmaf(dt, v.x, pre.x); }
  __device__ __forceinline__ void flush(long long) {}
};

// real lines [n_lines][n] -> half spectra [n_lines][n/2+1]
struct LfIoR2C {
  const float* in;
  float2* out;
  int n, hp;
  __device__ __forceinline__ float2 load_pair(long long pair, int idx) const {
    const float* a = in + 2 * pair * n;
    return make_float2(a[idx], a[n + idx]);
  }
  __device__ __forceinline__ void store_half(long long pair, int h, float2 A, float2 B) const {
    float2* o = out + 2 * pair * hp;
    o[h] = A;
    o[hp + h] = B;
  }
};
// half spectra -> y1 = y0 + dt * real lines   (solvers.py:63)
struct LfIoC2RUpdate {
  const float2* in;
  const float* y0;
  float* y1;
  int n, hp;
  float dt;
  __device__ __forceinline__ void load_half(long long pair, int h, float2& A, float2& B) const {
    const float2* a = in + 2 * pair * hp;
    A = a[h];
    B = a[hp + h];
  }
  __device__ __forceinline__ float2 pre_pair(long long pair, int idx) const {
    const long long o = 2 * pair * n + idx;
    return make_float2(y0[o], y0[o + n]);
  }
  __device__ __forceinline__ void store_pair(long long pair, int idx, float2 v, float2 pre) const {
    const long long o = 2 * pair * n + idx;
    y1[o] = fmaf(dt, v.x, pre.x);
    y1[o + n] = fmaf(dt, v.y, pre.y);
  }
};

}  // namespace pdeopt
