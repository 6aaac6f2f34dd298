// 256 x 256 complex FFT of ONE field distributed over a thread-block cluster of 4 CTAs (sm_100a),
// host-emulable: the building block of the kinetic Strang step of GPE2DTSControl on BASELINE config 3
// (replaces the jnp.fft.fftn / ifftn calls of StrangSplitting.step, solvers.py:107-114).
//
// The 512 KB field never leaves the cluster: CTA q holds 64 lines of 256 points (128 KB) in registers,
// 32 complex values per thread, either as a ROW slab (rows 64q..64q+63, all columns) or as a COLUMN
// slab.  A pass transforms the 64 local lines; the row <-> column transposes in between are remote
// stores into the peers' shared memory (distributed shared memory, st.shared::cluster): the last
// stage of a pass writes every value straight to the CTA and slot where the next pass will read it.
//
// One line = 8 threads (consecutive lanes of one warp) x 32 points:
//   spatial arrangement   thread (l, j) holds n = 8 n1 + j,            n1 = 0..31  -> x[n1]
//   S1   32-point DFT over n1 (in thread), twiddle w256^(j k1)
//   E1   exchange among the 8 threads of the line through the line's own 2 KB of shared memory
//        (same warp: __syncwarp, no CTA barrier)
//   S2   four 8-point DFTs over j                                        -> k = k1 + 32 k0
//   frequency arrangement thread (l, j') holds k = j' + 8 a + 32 k0,    a = 0..3, k0 = 0..7 -> x[8 a + k0]
// The inverse mirrors it.  Shared-memory layout of a slab: 8-byte slot(line, pos) = 256 line + (pos ^ g(line)),
// g(line) = ((line & 1) << 3) | (line & 6): every access pattern below (remote transposed stores included,
// whose bank conflicts are paid at the DESTINATION) hits 16 distinct banks per half-warp.
// tests/test_cfft256_host.py runs this header on the host (4 emulated CTAs x 512 threads) against numpy.
#pragma once
#include <stdint.h>

#include "regfft.cuh"

namespace pdeopt {
namespace cf {

constexpr int kN = 256;
constexpr int kCtas = 4;
constexpr int kLines = kN / kCtas;  // lines per CTA
constexpr int kThreadsC = 512;      // 8 threads per line
constexpr uint32_t kSlabBytes = kLines * kN * 8;  // 128 KB

PDEOPT_HD int g_of(int line) { return ((line & 1) << 3) | (line & 6); }
PDEOPT_HD uint32_t slot_bytes(int line, int pos) { return (uint32_t)((line * kN + (pos ^ g_of(line))) * 8); }
// exchange E1 inside a line's region: value (k1, j) at 32 j + (k1 ^ j ^ ((line & 1) << 3))
PDEOPT_HD uint32_t e1_bytes(int line, int k1, int j) { return (uint32_t)((line * kN + 32 * j + (k1 ^ j ^ ((line & 1) << 3))) * 8); }

// ---- memory access: device = shared-window addresses (+ mapa for peers); host = emulated slabs ----
#if defined(__CUDACC__)
__device__ __forceinline__ float2 lds(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t peer_addr(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void sts_peer(uint32_t mapped, float2 v) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(mapped), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void line_sync() { __syncwarp(); }
#define PDEOPT_CF_FN __device__ __forceinline__
struct Ctx {
  uint32_t base;           // this CTA's slab (shared-window byte address)
  uint32_t peer[kCtas];    // the same address mapped into every CTA of the cluster (own rank included)
  __device__ __forceinline__ float2 ld(uint32_t off) const { return lds(base + off); }
  __device__ __forceinline__ void st(uint32_t off, float2 v) const { sts(base + off, v); }
  __device__ __forceinline__ void st_to(int rank, uint32_t off, float2 v) const { sts_peer(peer[rank] + off, v); }
};
#else
#define PDEOPT_CF_FN inline
inline void line_sync() {}
struct Ctx {
  unsigned char* slabs[kCtas];  // emulated shared memory of the four CTAs
  int rank;
  float2 ld(uint32_t off) const { return *reinterpret_cast<const float2*>(slabs[rank] + off); }
  void st(uint32_t off, float2 v) const { *reinterpret_cast<float2*>(slabs[rank] + off) = v; }
  void st_to(int r, uint32_t off, float2 v) const { *reinterpret_cast<float2*>(slabs[r] + off) = v; }
};
#endif

// ---- loads of a line from the slab ------------------------------------------------------------------
// spatial arrangement, placed bit-reversed for the decimation-in-time S1: x[brev5(n1)] = line[8 n1 + j]
PDEOPT_CF_FN void load_spatial(const Ctx& c, int l, int j, float2 (&x)[32]) {
  static_for<0, 32>([&](auto nc) {
    constexpr int n1 = decltype(nc)::value;
    x[brev<5>(n1)] = c.ld(slot_bytes(l, 8 * n1 + j));
  });
}
// frequency arrangement in natural register order: x[8 a + k0] = line[j + 8 a + 32 k0]
PDEOPT_CF_FN void load_freq(const Ctx& c, int l, int j, float2 (&x)[32]) {
  static_for<0, 32>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    x[i] = c.ld(slot_bytes(l, j + 8 * (i >> 3) + 32 * (i & 7)));
  });
}

// ---- forward transform of a line: x[brev5(n1)] spatial in -> x[8 a + k0] frequency out ---------------
// tw: [8][32] float2, tw[j][k1] = w256^(j k1) (forward sign).  Two parts with the line's exchange in
// between (the host emulation runs each part for all threads of a line before the next one).
PDEOPT_CF_FN void line_fwd_a(const Ctx& c, const float2* __restrict__ tw, int l, int j, float2 (&x)[32]) {
  DitF<32, 1, false>::run(x);  // x[k1]
  const float2* t = tw + j * 32;
  static_for<1, 32>([&](auto kc) {
    constexpr int k1 = decltype(kc)::value;
    x[k1] = cmul(x[k1], t[k1]);
  });
  line_sync();  // every thread of the line has read its inputs from the line's region
  static_for<0, 32>([&](auto kc) {
    constexpr int k1 = decltype(kc)::value;
    c.st(e1_bytes(l, k1, j), x[k1]);
  });
}
PDEOPT_CF_FN void line_fwd_b(const Ctx& c, int l, int j, float2 (&x)[32]) {
  static_for<0, 4>([&](auto ac) {
    constexpr int a = decltype(ac)::value;
    static_for<0, 8>([&](auto jc) {
      constexpr int jj = decltype(jc)::value;
      x[8 * a + brev<3>(jj)] = c.ld(e1_bytes(l, j + 8 * a, jj));
    });
  });
  static_for<0, 4>([&](auto ac) { DitF<8, 1, false>::run(x + 8 * decltype(ac)::value); });
}
PDEOPT_CF_FN void line_fwd(const Ctx& c, const float2* __restrict__ tw, int l, int j, float2 (&x)[32]) {
  line_fwd_a(c, tw, l, j, x);
  line_sync();
  line_fwd_b(c, l, j, x);
}

// ---- inverse transform of a line: x[8 a + k0] frequency in -> x[n1] spatial out (times 256) ----------
PDEOPT_CF_FN void line_inv_a(const Ctx& c, int l, int j, float2 (&x)[32]) {
  // natural k0 order -> bit-reversed placement for the decimation-in-time inverse (register renaming)
  static_for<0, 4>([&](auto ac) {
    constexpr int a = decltype(ac)::value;
    float2 t1 = x[8 * a + 1], t3 = x[8 * a + 3];
    x[8 * a + 1] = x[8 * a + 4];
    x[8 * a + 4] = t1;
    x[8 * a + 3] = x[8 * a + 6];
    x[8 * a + 6] = t3;
    DitF<8, 1, true>::run(x + 8 * a);  // x[8 a + jj]: value for thread jj of the line, k1 = j + 8 a
  });
  line_sync();  // every thread of the line has read its inputs from the line's region
  static_for<0, 4>([&](auto ac) {
    constexpr int a = decltype(ac)::value;
    static_for<0, 8>([&](auto jc) {
      constexpr int jj = decltype(jc)::value;
      c.st(e1_bytes(l, j + 8 * a, jj), x[8 * a + jj]);
    });
  });
}
PDEOPT_CF_FN void line_inv_b(const Ctx& c, const float2* __restrict__ tw, int l, int j, float2 (&x)[32]) {
  const float2* t = tw + j * 32;
  static_for<0, 32>([&](auto kc) {
    constexpr int k1 = decltype(kc)::value;
    float2 v = c.ld(e1_bytes(l, k1, j));
    if constexpr (k1 > 0) v = cmulc(v, t[k1]);
    x[brev<5>(k1)] = v;
  });
  DitF<32, 1, true>::run(x);  // x[n1]
}
PDEOPT_CF_FN void line_inv(const Ctx& c, const float2* __restrict__ tw, int l, int j, float2 (&x)[32]) {
  line_inv_a(c, l, j, x);
  line_sync();
  line_inv_b(c, tw, l, j, x);
}

// ---- transposed stores: the row <-> column exchange across the cluster --------------------------------
// frequency arrangement of line `gl` (global line index 0..255) -> the slab that owns position k as a line:
// value k = j + 8 a + 32 k0 goes to CTA k >> 6, line k & 63, position gl
PDEOPT_CF_FN void store_transposed_from_freq(const Ctx& c, int gl, int j, const float2 (&x)[32]) {
  static_for<0, 32>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int a = i >> 3, k0 = i & 7;
    const int line = j + 8 * a + 32 * (k0 & 1);
    c.st_to(k0 >> 1, slot_bytes(line, gl), x[i]);
  });
}
// spatial arrangement of line `gl` -> value n = 8 n1 + j goes to CTA n >> 6 = n1 >> 3, line 8 (n1 & 7) + j
PDEOPT_CF_FN void store_transposed_from_spatial(const Ctx& c, int gl, int j, const float2 (&x)[32]) {
  static_for<0, 32>([&](auto nc) {
    constexpr int n1 = decltype(nc)::value;
    const int line = 8 * (n1 & 7) + j;
    c.st_to(n1 >> 3, slot_bytes(line, gl), x[n1]);
  });
}
// plain (non-transposed) store of the spatial arrangement into the own slab
PDEOPT_CF_FN void store_spatial(const Ctx& c, int l, int j, const float2 (&x)[32]) {
  static_for<0, 32>([&](auto nc) {
    constexpr int n1 = decltype(nc)::value;
    c.st(slot_bytes(l, 8 * n1 + j), x[n1]);
  });
}

}  // namespace cf
}  // namespace pdeopt
