// 256 x 256 complex FFT of ONE field distributed over a thread-block cluster of 8 (or 4) CTAs (sm_100a),
// host-emulable: the building block of the kinetic Strang step of GPE2DTSControl on BASELINE config 3
// (replaces the jnp.fft.fftn / ifftn calls of StrangSplitting.step, solvers.py:107-114).
//
// The 512 KB field never leaves the cluster: CTA q holds 64 lines of 256 points (128 KB) in registers,
// 32 complex values per thread, either as a ROW slab (rows 64q..64q+63, all columns) or as a COLUMN
// slab.  A pass transforms the 64 local lines; the row <-> column transposes in between are remote
// stores into the peers' shared memory (distributed shared memory, st.shared::cluster): the last
// stage of a pass writes every value straight to the CTA and slot where the next pass will read it.
//
// One line = 8 threads x 32 points; the LANES of a warp are 32 different lines (thread t: line
// l = (t & 31) + 32 ((t >> 5) & 1), j = t >> 6), so that a transposed store of a warp is 32 consecutive
// positions of ONE destination line: 256 contiguous bytes per st.shared::cluster instruction (scattered
// 8-byte remote stores run at a few bytes per clock):
//   spatial arrangement   thread (l, j) holds n = 8 n1 + j,            n1 = 0..31  -> x[n1]
//   S1   32-point DFT over n1 (in thread), twiddle w256^(j k1)
//   E1   exchange among the 8 threads of the line through the line's own 2 KB of shared memory
//        (they sit in different warps: CTA barriers around it)
//   S2   four 8-point DFTs over j                                        -> k = k1 + 32 k0
//   frequency arrangement thread (l, j') holds k = j' + 8 a + 32 k0,    a = 0..3, k0 = 0..7 -> x[8 a + k0]
// The inverse mirrors it.  Shared-memory layout of a slab: 8-byte slot(line, pos) = 256 line + (pos ^ (line & 15)):
// every access pattern below hits 16 distinct banks per half-warp (16 consecutive lines).
// tests/test_cfft256_host.py runs this header on the host (4 emulated CTAs x 512 threads) against numpy.
#pragma once
#include <stdint.h>

#include "regfft.cuh"

namespace pdeopt {
namespace cf {

constexpr int kN = 256;
// Cluster shape.  4 CTAs x 512 threads (64 lines, 128 KB per CTA, one CTA per SM) is the default.  PDEOPT_CF_CTAS=8 builds
// 8 CTAs x 256 threads (32 lines, 64 KB per CTA): two CTAs of DIFFERENT clusters then share an SM and one environment's
// exchanges and barriers overlap the other's arithmetic — measured: 30 environments resident take 51 us per step against 32 us for
// 15 (+27 % per SM) — but only 15 clusters of 8 fit the GPCs of a B200 per layer (120 of 148 SMs; 37 clusters of 4 use all of
// them), so both shapes end at 630 k env-steps/s for 128 environments.  Both are checked by the host emulation.
#ifndef PDEOPT_CF_CTAS
#define PDEOPT_CF_CTAS 4
#endif
constexpr int kCtas = PDEOPT_CF_CTAS;
static_assert(kCtas == 4 || kCtas == 8, "cluster of 4 or 8 CTAs");
constexpr int kLines = kN / kCtas;  // lines per CTA (64 or 32)
constexpr int kThreadsC = 8 * kLines;  // 8 threads per line
constexpr uint32_t kSlabBytes = kLines * kN * 8;  // 128 KB or 64 KB
constexpr int kLineGroups = kLines / 8;  // destination lines j + 8 m, m < kLineGroups

PDEOPT_HD int g_of(int line) { return line & 15; }
PDEOPT_HD uint32_t slot_bytes(int line, int pos) { return (uint32_t)((line * kN + (pos ^ g_of(line))) * 8); }
// exchange E1 inside a line's region: value (k1, j) at 32 j + (k1 ^ (line & 15))
PDEOPT_HD uint32_t e1_bytes(int line, int k1, int j) { return (uint32_t)((line * kN + 32 * j + (k1 ^ (line & 15))) * 8); }
PDEOPT_HD int thread_line(int t) { return t % kLines; }  // the lanes of a warp are 32 consecutive lines
PDEOPT_HD int thread_j(int t) { return t / kLines; }

// ---- memory access: device = shared-window addresses (+ mapa for peers); host = emulated slabs ----
// All accesses are [register + compile-time immediate]; the slab is 2048-byte aligned so that the swizzle
// XORs (bits 3..10) commute with the base address.
#if defined(__CUDACC__)
#define PDEOPT_CF_FN __device__ __forceinline__
__device__ __forceinline__ void line_sync() { __syncthreads(); }
struct Ctx {
  uint32_t base;         // this CTA's slab (shared-window byte address, 2048-byte aligned)
  uint32_t peer[kCtas];  // the same address mapped into every CTA of the cluster (own rank: the local address)
  int self;              // this CTA's rank in the cluster
  template <int OFF>
  __device__ __forceinline__ float2 ld(uint32_t a) const {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(a), "n"(OFF));
    return v;
  }
  template <int OFF>
  __device__ __forceinline__ void st(uint32_t a, float2 v) const {
    asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "f"(v.x), "f"(v.y) : "memory");
  }
  // `a` is already an address inside the peer's window (LineMap::Tp); the quarter of the transposed data
  // that stays in this CTA takes the ordinary shared-memory path (Tp[self] is a local address)
  template <int OFF>
  __device__ __forceinline__ void st_to(int rank, uint32_t a, float2 v) const {
    if (rank == self) {
      asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "f"(v.x), "f"(v.y) : "memory");
    } else {
      asm volatile("st.shared::cluster.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "f"(v.x), "f"(v.y) : "memory");
    }
  }
  __device__ __forceinline__ uint32_t local(uint32_t off) const { return base + off; }
  __device__ __forceinline__ uint32_t remote(int rank, uint32_t off) const { return peer[rank] + off; }
};
__device__ __forceinline__ uint32_t peer_addr(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
#else
#define PDEOPT_CF_FN inline
inline void line_sync() {}
struct Ctx {
  unsigned char* slabs[kCtas];  // emulated shared memory of the four CTAs
  int rank;
  template <int OFF>
  float2 ld(uint32_t a) const { return *reinterpret_cast<const float2*>(slabs[rank] + a + OFF); }
  template <int OFF>
  void st(uint32_t a, float2 v) const { *reinterpret_cast<float2*>(slabs[rank] + a + OFF) = v; }
  template <int OFF>
  void st_to(int r, uint32_t a, float2 v) const { *reinterpret_cast<float2*>(slabs[r] + a + OFF) = v; }
  uint32_t local(uint32_t off) const { return off; }
  uint32_t remote(int, uint32_t off) const { return off; }
};
#endif

// Per-thread address bases: every slab address of a thread is one of these, XOR a one-bit constant, plus a
// compile-time immediate (the swizzle occupies address bits 3..6, the immediates bits >= 7):
//   slot(l, 8 i + j) / slot(l, j + 8 a + 32 k0) with i = a + 4 k0   = (B ^ 64 (i & 1)) + 128 (i >> 1)
//   e1(l, k1, j)   [own j]                                          = (E ^ 8 (k1 & 15)) + 128 (k1 >> 4)
//   e1(l, j + 8 a, jj) [thread jj's value]                          = (B ^ 64 (a & 1)) + 128 (a >> 1) + 256 jj
//   slot(j + 8 m, gl) in CTA r                                      = (Tp[r] ^ 64 (m & 1)) + 16384 m
// with B = slab + 2048 l + 8 (j ^ (l & 15)), E = slab + 2048 l + 256 j + 8 (l & 15), Tp[r] = peer r's slab + 2048 j + 8 (gl ^ j).
struct LineMap {
  uint32_t B, E;
#if defined(__CUDACC__)
  // The kCtas transposed-store bases live in shared memory (one row per thread, written once): eight more live
  // registers would be spilled to local memory, whose L1 lines every cluster barrier invalidates.
  uint32_t tp_row;  // shared-window address of this thread's row of kCtas bases
  __device__ __forceinline__ LineMap(const Ctx& c, int l, int j, int gl, uint32_t tp_table, int tid) {
    B = c.local((uint32_t)(2048 * l + 8 * (j ^ (l & 15))));
    E = c.local((uint32_t)(2048 * l + 256 * j + 8 * (l & 15)));
    tp_row = tp_table + (uint32_t)tid * (uint32_t)(kCtas * 4);
    for (int r = 0; r < kCtas; ++r) {
      const uint32_t v = c.remote(r, (uint32_t)(2048 * j + 8 * (gl ^ j)));
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(tp_row + 4u * (uint32_t)r), "r"(v) : "memory");
    }
  }
  __device__ __forceinline__ uint32_t tp(const Ctx&, int r) const {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tp_row + 4u * (uint32_t)r));
    return v;
  }
#else
  uint32_t Tp[kCtas];
  LineMap(const Ctx& c, int l, int j, int gl) {
    B = c.local((uint32_t)(2048 * l + 8 * (j ^ (l & 15))));
    E = c.local((uint32_t)(2048 * l + 256 * j + 8 * (l & 15)));
    for (int r = 0; r < kCtas; ++r) Tp[r] = c.remote(r, (uint32_t)(2048 * j + 8 * (gl ^ j)));
  }
  uint32_t tp(const Ctx&, int r) const { return Tp[r]; }
#endif
};

// ---- loads of a line from the slab ------------------------------------------------------------------
// spatial arrangement, placed bit-reversed for the decimation-in-time S1: x[brev5(n1)] = line[8 n1 + j]
PDEOPT_CF_FN void load_spatial(const Ctx& c, const LineMap& m, float2 (&x)[32]) {
  static_for<0, 32>([&](auto nc) {
    constexpr int n1 = decltype(nc)::value;
    x[brev<5>(n1)] = c.template ld<128 * (n1 >> 1)>(m.B ^ (uint32_t)(64 * (n1 & 1)));
  });
}
// frequency arrangement in natural register order: x[8 a + k0] = line[j + 8 a + 32 k0]
PDEOPT_CF_FN void load_freq(const Ctx& c, const LineMap& m, float2 (&x)[32]) {
  static_for<0, 32>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int idx = (i >> 3) + 4 * (i & 7);
    x[i] = c.template ld<128 * (idx >> 1)>(m.B ^ (uint32_t)(64 * (idx & 1)));
  });
}
// plain store of the spatial arrangement (natural n1 order) into the own slab
PDEOPT_CF_FN void store_spatial(const Ctx& c, const LineMap& m, const float2 (&x)[32]) {
  static_for<0, 32>([&](auto nc) {
    constexpr int n1 = decltype(nc)::value;
    c.template st<128 * (n1 >> 1)>(m.B ^ (uint32_t)(64 * (n1 & 1)), x[n1]);
  });
}

// ---- forward transform of a line: x[brev5(n1)] spatial in -> x[8 a + k0] frequency out ---------------
// tw: [8][32] float2, tw[j][k1] = w256^(j k1) (forward sign), `t` = tw + 32 j.  Two parts with the line's
// exchange in between (the host emulation runs each part for all threads of a line before the next one).
PDEOPT_CF_FN void line_fwd_a(const Ctx& c, const float2* __restrict__ t, const LineMap& m, float2 (&x)[32]) {
  DitF<32, 1, false>::run(x);  // x[k1]
  line_sync();  // every thread of the line has read its inputs from the line's region
  // twiddle and store element by element (keeps at most a few twiddles live in registers)
  static_for<0, 32>([&](auto kc) {
    constexpr int k1 = decltype(kc)::value;
    float2 v = x[k1];
    if constexpr (k1 > 0) v = cmul(v, t[k1]);
    c.template st<128 * (k1 >> 4)>(m.E ^ (uint32_t)(8 * (k1 & 15)), v);
  });
}
PDEOPT_CF_FN void line_fwd_b(const Ctx& c, const LineMap& m, float2 (&x)[32]) {
  static_for<0, 4>([&](auto ac) {
    constexpr int a = decltype(ac)::value;
    static_for<0, 8>([&](auto jc) {
      constexpr int jj = decltype(jc)::value;
      x[8 * a + brev<3>(jj)] = c.template ld<128 * (a >> 1) + 256 * jj>(m.B ^ (uint32_t)(64 * (a & 1)));
    });
  });
  static_for<0, 4>([&](auto ac) { DitF<8, 1, false>::run(x + 8 * decltype(ac)::value); });
}
PDEOPT_CF_FN void line_fwd(const Ctx& c, const float2* __restrict__ t, const LineMap& m, float2 (&x)[32]) {
  line_fwd_a(c, t, m, x);
  line_sync();
  line_fwd_b(c, m, x);
}

// ---- inverse transform of a line: x[8 a + k0] frequency in -> x[n1] spatial out (times 256) ----------
PDEOPT_CF_FN void line_inv_a(const Ctx& c, const LineMap& m, float2 (&x)[32]) {
  line_sync();  // every thread of the line has read its inputs from the line's region
  static_for<0, 4>([&](auto ac) {
    constexpr int a = decltype(ac)::value;
    // natural k0 order -> bit-reversed placement for the decimation-in-time inverse (register renaming)
    float2 t1 = x[8 * a + 1], t3 = x[8 * a + 3];
    x[8 * a + 1] = x[8 * a + 4];
    x[8 * a + 4] = t1;
    x[8 * a + 3] = x[8 * a + 6];
    x[8 * a + 6] = t3;
    DitF<8, 1, true>::run(x + 8 * a);  // x[8 a + jj]: value for thread jj of the line, k1 = j + 8 a
    static_for<0, 8>([&](auto jc) {
      constexpr int jj = decltype(jc)::value;
      c.template st<128 * (a >> 1) + 256 * jj>(m.B ^ (uint32_t)(64 * (a & 1)), x[8 * a + jj]);
    });
  });
}
PDEOPT_CF_FN void line_inv_b(const Ctx& c, const float2* __restrict__ t, const LineMap& m, float2 (&x)[32]) {
  static_for<0, 32>([&](auto kc) {
    constexpr int k1 = decltype(kc)::value;
    float2 v = c.template ld<128 * (k1 >> 4)>(m.E ^ (uint32_t)(8 * (k1 & 15)));
    if constexpr (k1 > 0) v = cmulc(v, t[k1]);
    x[brev<5>(k1)] = v;
  });
  DitF<32, 1, true>::run(x);  // x[n1]
}
PDEOPT_CF_FN void line_inv(const Ctx& c, const float2* __restrict__ t, const LineMap& m, float2 (&x)[32]) {
  line_inv_a(c, m, x);
  line_sync();
  line_inv_b(c, t, m, x);
}

// ---- transposed stores: the row <-> column exchange across the cluster --------------------------------
// frequency arrangement of line gl -> the slab that owns position k as a line: value k = j + 8 a + 32 k0
// goes to CTA k >> 6, line k & 63 = j + 8 (a + 4 (k0 & 1)), position gl
PDEOPT_CF_FN void store_transposed_from_freq(const Ctx& c, const LineMap& m, const float2 (&x)[32]) {
  static_for<0, 32>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int a = i >> 3, k0 = i & 7;
    constexpr int k = 8 * a + 32 * k0;                       // position (without j): destination line = (j + k) % kLines
    constexpr int rank = k / kLines, mm = (k % kLines) / 8;  // destination CTA and line group
    c.template st_to<16384 * mm>(rank, m.tp(c, rank) ^ (uint32_t)(64 * (mm & 1)), x[i]);
  });
}
// spatial arrangement of line gl -> value n = 8 n1 + j goes to CTA n / kLines, line 8 (n1 % kLineGroups) + j, position gl
PDEOPT_CF_FN void store_transposed_from_spatial(const Ctx& c, const LineMap& m, const float2 (&x)[32]) {
  static_for<0, 32>([&](auto nc) {
    constexpr int n1 = decltype(nc)::value;
    constexpr int rank = n1 / kLineGroups, mm = n1 % kLineGroups;
    c.template st_to<16384 * mm>(rank, m.tp(c, rank) ^ (uint32_t)(64 * (mm & 1)), x[n1]);
  });
}

}  // namespace cf
}  // namespace pdeopt
