// Spectral filter of ONE real 128x128 field held by one 256-thread CTA (sm_100a), host-emulable.
//
//   g = Re ifft2( fft2(f) * M ),   M real and even in both wavenumbers  (solvers.py:62-63 with
//   M = 1 / (1 + A dt sigma); replaces the jnp.fft.fftn / ifftn pair of the reference)
//
// The real field is packed along the columns into a 128 x 64 complex field
//   z[r][m] = f[r][2m] + i f[r][2m+1],
// transformed by a 128 x 64 complex FFT (13 index bits, three register passes, 32 complex values
// per thread), filtered in closed form
//   ZG[kr][km] = a[kr][km] Z[kr][km] + i b[kr][km] conj(Z[-kr][-km])
//   a = P - Q sin(th), b = Q cos(th), th = 2 pi km / 128, P/Q = (M[kr][km] +/- M[kr][km+64]) / 2
// (no Hermitian untangling pass; the spectrum is dealt to the threads so that every (k, -k) pair
// lives in ONE thread) and transformed back: zg = ifft2(ZG) holds g[r][2m] + i g[r][2m+1].
// tools/realfft_filter_model.py is the NumPy model of the algebra, the thread maps and the two
// exchange layouts (bijective, bank-conflict free); tests/test_rfft128_host.py runs THIS header on
// the host, 256 emulated threads per barrier phase, against numpy.fft.
//
// Passes (forward; the inverse mirrors them):
//   A  radix-32 over the high 5 bits of m            thread (r, m0),      m = 2 n + m0
//   B  twiddle w64^(m0 k1m), radix-2 over m0, radix-16 over the high 4 bits of r,
//      twiddle w128^(n2r k1r)                         thread (k1m, n2r),   r = 8 n1r + n2r
//   C  4 x radix-8 over n2r                           thread (k1r class j, km class c):
//      k1r in {j, 16-j} ({0, 8} for j = 0), km in {c, 64-c} ({0, 32} for c = 0), kr = k1r + 16 k2r
// Half the exchange traffic per field of the pair kernel's 128 x 128 complex transform, and two
// such CTAs (two independent fields, in different phases) share one SM.
#pragma once
#include <stdint.h>

#include "regfft.cuh"

namespace pdeopt {
namespace rf {

constexpr int kRows = 128;      // field rows
constexpr int kCols = 128;      // field columns (real)
constexpr int kH = 64;          // complex columns
constexpr int kThreadsR = 256;  // threads per CTA
constexpr int kTabDim = 65;     // folded symbol table is [65][65]
constexpr int kTRows = 65;      // filter table rows (|kr| = 0..64)
constexpr int kTCols = 32;      // filter table columns (km class c = 0..31)
constexpr uint32_t kWBytes = kRows * kH * 8;

// ---- shared-memory access: [base register (+ XOR of low bits)] + compile-time immediate --------
// Device: 32-bit shared-window byte addresses.  Host emulation: byte offsets into g_emul.
#if defined(__CUDACC__)
template <int OFF>
__device__ __forceinline__ float2 ld2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ void st2(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "f"(v.x), "f"(v.y) : "memory");
}
template <int OFF>
__device__ __forceinline__ float4 ld4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF));
  return v;
}
#define PDEOPT_RF_FN __device__ __forceinline__
#else
#ifndef PDEOPT_HOST_FLOAT4
#define PDEOPT_HOST_FLOAT4
struct float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float a, float b, float c, float d) {
  float4 r;
  r.x = a; r.y = b; r.z = c; r.w = d;
  return r;
}
#endif
static unsigned char* g_emul = nullptr;  // host emulation of the CTA's shared memory
template <int OFF>
inline float2 ld2(uint32_t a) {
  return *reinterpret_cast<const float2*>(g_emul + (int64_t)a + OFF);
}
template <int OFF>
inline void st2(uint32_t a, float2 v) {
  *reinterpret_cast<float2*>(g_emul + (int64_t)a + OFF) = v;
}
template <int OFF>
inline float4 ld4(uint32_t a) {
  return *reinterpret_cast<const float4*>(g_emul + (int64_t)a + OFF);
}
#define PDEOPT_RF_FN inline
#endif

// natural (spatial) layout of z: 8-byte slot of (r, m)
PDEOPT_HD int nat_slot(int r, int m) { return r * kH + (m ^ ((r & 7) << 1)); }

// Per-thread addressing state of the three passes.  `w` = byte address of the 64 KB field buffer
// (1024-byte aligned), `t` = byte address of the filter table, `tid` = thread index in the CTA.
struct RFft {
  uint32_t nb;    // natural layout, pass-A map: element n = lo + 8 hi at (nb ^ lo << 4) + 128 hi
  // exchange layout (one layout for both exchanges, so passes B and C work in place and every
  // exchange costs one barrier): 8-byte slot of (k1m, R = 8 hi + n2r, q)
  //   ex_slot = 256 k1m + 16 hi + (((n2r << 1) | q) ^ (k1m & 15))
  // with (R, q) = (r, m0) between A and B and (8 k1r + n2r, k2m) between B and C.  Bank-conflict free:
  // the 16 lanes of a half-warp differ in (n2r, q) [pass A], (n2r, k1m[0]) [pass B] or k1m[3:0] [pass C].
  uint32_t ea;    // pass-A side: k1m at (ea ^ 8 (k1m & 15)) + 2048 k1m
  uint32_t bb;    // pass-B side: (q, hi) at (bb ^ 8 q) + 128 hi
  uint32_t d[4];  // pass-C side: slot s = 2 i + jj, n2r at d[s] ^ (n2r << 4)
  uint32_t tlo, thi;  // filter-table rows j + 16 k2r (k2r <= 3) and 64 - j - 16 (k2r - 4) (k2r >= 4)
  int b_n2r, b_k1m;   // pass-B indices (inter-pass twiddles)
  bool lane0, warp0;  // km class 0 / k1r class 0: self-conjugate classes

  PDEOPT_HD RFft(uint32_t w, uint32_t t, int tid) {
    {  // pass A: thread (r, m0)
      const int m0 = tid & 1, r = tid >> 1;
      nb = w + (uint32_t)(r * 512 + 8 * m0 + 16 * (r & 7));
      ea = w + (uint32_t)((r >> 3) * 128 + 8 * (((r & 7) << 1) | m0));
    }
    {  // pass B: thread (k1m, n2r); low four thread bits = (n2r, k1m[0])
      const int k1m0 = tid & 1;
      b_n2r = (tid >> 1) & 7;
      b_k1m = k1m0 | ((tid >> 4) << 1);
      bb = w + (uint32_t)(b_k1m * 2048 + 8 * ((b_n2r << 1) ^ (b_k1m & 15)));
    }
    {  // pass C: warp = k1r class j, lane = km class c
      const int c = tid & 31, j = tid >> 5;
      lane0 = c == 0;
      warp0 = j == 0;
      const int k1r[2] = {j, j == 0 ? 8 : 16 - j};
      const int k1m[2] = {c, (32 - c) & 31};
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int i = 0; i < 2; ++i)
        for (int jj = 0; jj < 2; ++jj)
          d[2 * i + jj] = w + (uint32_t)(k1m[jj] * 2048 + k1r[i] * 128 + 8 * (jj ^ (k1m[jj] & 15)));
      tlo = t + (uint32_t)((j * kTCols + c) * 16);
      thi = t + (uint32_t)(((64 - j) * kTCols + c) * 16);
    }
  }
};

// ---- natural layout <-> pass-A registers: x[n] <-> z[r][2 n + m0] --------------------------------
// All transforms are decimation in time (six-slot butterflies, regfft.cuh DitF), which consume their
// input in bit-reversed register order: BREV = true places element n in register brev5(n) (free: the
// register indices are compile-time).  scatter_nat / gather_nat<false> use the natural order.
template <bool BREV = false>
PDEOPT_RF_FN void gather_nat(const RFft& F, float2 (&x)[32]) {
  static_for<0, 8>([&](auto lc) {
    constexpr int lo = decltype(lc)::value;
    const uint32_t a = F.nb ^ (uint32_t)(lo << 4);
    static_for<0, 4>([&](auto hc) {
      constexpr int hi = decltype(hc)::value;
      constexpr int n = hi * 8 + lo;
      x[BREV ? brev<5>(n) : n] = ld2<hi * 128>(a);
    });
  });
}
PDEOPT_RF_FN void scatter_nat(const RFft& F, const float2 (&x)[32]) {
  static_for<0, 8>([&](auto lc) {
    constexpr int lo = decltype(lc)::value;
    const uint32_t a = F.nb ^ (uint32_t)(lo << 4);
    static_for<0, 4>([&](auto hc) {
      constexpr int hi = decltype(hc)::value;
      st2<hi * 128>(a, x[hi * 8 + lo]);
    });
  });
}

// ---- pass A ------------------------------------------------------------------------------------
// forward: 32-point DFT over n (x[brev5(n)] in, from gather_nat<true>), k1m = position out
PDEOPT_RF_FN void passA_fwd(const RFft& F, float2 (&x)[32]) {
  DitF<32, 1, false>::run(x);
  static_for<0, 32>([&](auto kc) {
    constexpr int k1m = decltype(kc)::value;
    st2<k1m * 2048>(F.ea ^ (uint32_t)((k1m & 15) * 8), x[k1m]);
  });
}
// inverse: loads the A -> B layout, returns 8192 * zg in the pass-A arrangement (natural n)
PDEOPT_RF_FN void passA_inv(const RFft& F, float2 (&x)[32]) {
  static_for<0, 32>([&](auto kc) {
    constexpr int k1m = decltype(kc)::value;
    x[brev<5>(k1m)] = ld2<k1m * 2048>(F.ea ^ (uint32_t)((k1m & 15) * 8));
  });
  DitF<32, 1, true>::run(x);
}

// ---- pass B ------------------------------------------------------------------------------------
// twb: [16][8] float2, twb[k1r][n2r] = w128^(n2r k1r) (forward sign; the eight entries a warp reads
// at once are contiguous: one wavefront); tw64: [32] float2 = w64^k1m
PDEOPT_RF_FN void passB_fwd(const RFft& F, const float2* __restrict__ twb, const float2* __restrict__ tw64,
                            float2 (&x)[32]) {
  static_for<0, 2>([&](auto mc) {
    constexpr int m0 = decltype(mc)::value;
    const uint32_t a = F.bb ^ (uint32_t)(m0 * 8);
    static_for<0, 16>([&](auto nc) {
      constexpr int n1r = decltype(nc)::value;
      x[m0 * 16 + brev<4>(n1r)] = ld2<n1r * 128>(a);
    });
  });
  {
    // radix 2 over m0 with the twiddle fused: p = a + w b, q = 2a - p
    const float2 w = tw64[F.b_k1m];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int n = 0; n < 16; ++n) {
      const float2 a = x[n], b = x[16 + n];
      const float2 p = cmac(a, b, w.x, w.y);
      x[n] = p;
      x[16 + n] = twice_minus(a, p);
    }
  }
  DitF<16, 1, false>::run(x);
  DitF<16, 1, false>::run(x + 16);
  {
    const float2* tw = twb + F.b_n2r;
    static_for<1, 16>([&](auto pc) {
      constexpr int k1r = decltype(pc)::value;
      const float2 w = tw[k1r * 8];
      x[k1r] = cmul(x[k1r], w);
      x[16 + k1r] = cmul(x[16 + k1r], w);
    });
  }
  static_for<0, 2>([&](auto kc) {
    constexpr int k2m = decltype(kc)::value;
    const uint32_t a = F.bb ^ (uint32_t)(k2m * 8);
    static_for<0, 16>([&](auto pc) {
      constexpr int k1r = decltype(pc)::value;
      st2<k1r * 128>(a, x[k2m * 16 + k1r]);
    });
  });
}
PDEOPT_RF_FN void passB_inv(const RFft& F, const float2* __restrict__ twb, const float2* __restrict__ tw64,
                            float2 (&x)[32]) {
  static_for<0, 2>([&](auto kc) {
    constexpr int k2m = decltype(kc)::value;
    const uint32_t a = F.bb ^ (uint32_t)(k2m * 8);
    static_for<0, 16>([&](auto pc) {
      constexpr int p = decltype(pc)::value;
      x[k2m * 16 + p] = ld2<brev<4>(p) * 128>(a);
    });
  });
  {
    const float2* tw = twb + F.b_n2r;
    static_for<1, 16>([&](auto pc) {
      constexpr int p = decltype(pc)::value;
      const float2 w = tw[brev<4>(p) * 8];
      x[p] = cmulc(x[p], w);
      x[16 + p] = cmulc(x[16 + p], w);
    });
  }
  DitF<16, 1, true>::run(x);
  DitF<16, 1, true>::run(x + 16);
  {
    const float2 w = tw64[F.b_k1m];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int n = 0; n < 16; ++n) {
      const float2 a = x[n], b = x[16 + n];
      x[n] = cadd(a, b);
      x[16 + n] = cmulc(csub(a, b), w);
    }
  }
  static_for<0, 2>([&](auto mc) {
    constexpr int m0 = decltype(mc)::value;
    const uint32_t a = F.bb ^ (uint32_t)(m0 * 8);
    static_for<0, 16>([&](auto nc) {
      constexpr int n1r = decltype(nc)::value;
      st2<n1r * 128>(a, x[m0 * 16 + n1r]);
    });
  });
}

// ---- pass C + closed-form filter ------------------------------------------------------------------
PDEOPT_RF_FN void passC_load(const RFft& F, float2 (&x)[32]) {
  static_for<0, 4>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    static_for<0, 8>([&](auto nc) {
      constexpr int n2r = decltype(nc)::value;
      x[s * 8 + n2r] = ld2<0>(F.d[s] ^ (uint32_t)(n2r << 4));
    });
  });
}
PDEOPT_RF_FN void passC_store(const RFft& F, const float2 (&x)[32]) {
  static_for<0, 4>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    static_for<0, 8>([&](auto nc) {
      constexpr int n2r = decltype(nc)::value;
      st2<0>(F.d[s] ^ (uint32_t)(n2r << 4), x[s * 8 + n2r]);
    });
  });
}

PDEOPT_HD float selp(bool c, float a, float b) { return c ? a : b; }
// a Z + i b conj(P) = (a Z.x + b P.y, a Z.y + b P.x): two packed instructions
PDEOPT_HD float2 filt(float a, float2 Z, float b, float2 P) {
  return fma2(make_float2(b, b), make_float2(P.y, P.x), mul2(make_float2(a, a), Z));
}

// Two mutually conjugate positions: slot row IA / position PA and slot row IB / position PB.  Elements
//   X = x[(IA,0),PA] (km = c),  U = x[(IA,1),PA] (km = 64 - c),  V = x[(IB,0),PB],  Y = x[(IB,1),PB];
// partners X <-> Y and U <-> V, except in km class 0 (lane 0: km in {0, 32} are their own negatives)
// where X <-> V and the coefficients b' of U and Y vanish.  t = (a, a', b, b') of this |kr| row.
template <int IA, int PA, int IB, int PB>
PDEOPT_RF_FN void filter_pair(float2 (&x)[32], const float4 t, bool lane0) {
  float2& X = x[(2 * IA + 0) * 8 + PA];
  float2& U = x[(2 * IA + 1) * 8 + PA];
  float2& V = x[(2 * IB + 0) * 8 + PB];
  float2& Y = x[(2 * IB + 1) * 8 + PB];
  const float2 pX = make_float2(selp(lane0, V.x, Y.x), selp(lane0, V.y, Y.y));
  const float2 pV = make_float2(selp(lane0, X.x, U.x), selp(lane0, X.y, U.y));
  const float2 nX = filt(t.x, X, t.z, pX);
  const float2 nV = filt(t.x, V, t.z, pV);
  const float2 nY = filt(t.y, Y, t.w, X);
  const float2 nU = filt(t.y, U, t.w, V);
  X = nX;
  V = nV;
  Y = nY;
  U = nU;
}
// A position whose row index is its own negative (kr in {0, 64}): X <-> U (lane 0: X <-> X).
template <int I, int P>
PDEOPT_RF_FN void filter_self(float2 (&x)[32], const float4 t, bool lane0) {
  float2& X = x[(2 * I + 0) * 8 + P];
  float2& U = x[(2 * I + 1) * 8 + P];
  const float2 pX = make_float2(selp(lane0, X.x, U.x), selp(lane0, X.y, U.y));
  const float2 nX = filt(t.x, X, t.z, pX);
  const float2 nU = filt(t.y, U, t.w, X);
  X = nX;
  U = nU;
}

// forward radix-8 passes, filter, inverse radix-8 passes; in place on the B -> C layout
PDEOPT_RF_FN void passC_filter(const RFft& F, float2 (&x)[32]) {
  passC_load(F, x);
  static_for<0, 4>([&](auto sc) { Dif<8, 1, false>::run(x + 8 * decltype(sc)::value); });
  // position p of a slot holds k2r = brev3(p); kr = k1r + 16 k2r
  if (!F.warp0) {
    // k1r in {j, 16 - j}: -(j + 16 k2r) = (16 - j) + 16 (7 - k2r), i.e. position 7 - p of the other row
    static_for<0, 8>([&](auto kc) {
      constexpr int k2r = decltype(kc)::value;
      constexpr int p = brev<3>(k2r);
      float4 t;
      if constexpr (k2r <= 3) {
        t = ld4<k2r * 8192>(F.tlo);
      } else {
        t = ld4<-(k2r - 4) * 8192>(F.thi);
      }
      filter_pair<0, p, 1, 7 - p>(x, t, F.lane0);
    });
  } else {
    // k1r = 0: -(16 k2r) = 16 ((8 - k2r) & 7), same row; rows |kr| = 0, 16, 32, 48, 64
    filter_self<0, brev<3>(0)>(x, ld4<0>(F.tlo), F.lane0);
    filter_self<0, brev<3>(4)>(x, ld4<4 * 8192>(F.tlo), F.lane0);
    static_for<1, 4>([&](auto kc) {
      constexpr int k2r = decltype(kc)::value;
      filter_pair<0, brev<3>(k2r), 0, brev<3>(8 - k2r)>(x, ld4<k2r * 8192>(F.tlo), F.lane0);
    });
    // k1r = 8: -(8 + 16 k2r) = 8 + 16 (7 - k2r), same row; rows |kr| = 8, 24, 40, 56
    static_for<0, 4>([&](auto kc) {
      constexpr int k2r = decltype(kc)::value;
      filter_pair<1, brev<3>(k2r), 1, 7 - brev<3>(k2r)>(x, ld4<4096 + k2r * 8192>(F.tlo), F.lane0);
    });
  }
  static_for<0, 4>([&](auto sc) { DitF<8, 1, true>::run(x + 8 * decltype(sc)::value); });
  passC_store(F, x);
}

// ---- filter table -------------------------------------------------------------------------------
// T[fr][c] = (a, a', b, b') / 8192 for |kr| = fr and km class c, from the folded symbol table
// tab[fr][fc] = A sigma(|kr|, |kc|) (kc = original column wavenumber index 0..127, folded):
//   M1 = M[fr][c], M2 = M[fr][c + 64] = M[fr][fold 64 - c];  a = P - Q sin, a' = P + Q sin (the class
//   member km = 64 - c), b = b' = Q cos;  class 0: member km = 32 has a' = M[fr][32], b' = 0.
// sc[c] = (cos, sin)(2 pi c / 128).
PDEOPT_HD float4 filter_entry(const float* __restrict__ tab, const float2* __restrict__ sc, int fr, int c, float dt) {
  const float inv = 1.0f / 8192.0f;
#if defined(__CUDA_ARCH__)
  const float m1 = __fdividef(inv, fmaf(dt, tab[fr * kTabDim + c], 1.0f));
  const float m2 = __fdividef(inv, fmaf(dt, tab[fr * kTabDim + 64 - c], 1.0f));
#else
  const float m1 = inv / fmaf(dt, tab[fr * kTabDim + c], 1.0f);
  const float m2 = inv / fmaf(dt, tab[fr * kTabDim + 64 - c], 1.0f);
#endif
  const float P = 0.5f * (m1 + m2), Q = 0.5f * (m1 - m2);
  const float2 cs = sc[c];
  float4 t;
  t.x = fmaf(-Q, cs.y, P);
  t.y = fmaf(Q, cs.y, P);
  t.z = Q * cs.x;
  t.w = t.z;
  if (c == 0) {
#if defined(__CUDA_ARCH__)
    t.y = __fdividef(inv, fmaf(dt, tab[fr * kTabDim + 32], 1.0f));
#else
    t.y = inv / fmaf(dt, tab[fr * kTabDim + 32], 1.0f);
#endif
    t.w = 0.0f;
  }
  return t;
}

}  // namespace rf
}  // namespace pdeopt
