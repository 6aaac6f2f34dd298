// Complex 128x128 FFT held in registers and shared memory by one 512-thread CTA (sm_100a).
//
// 14 index bits in three register passes (32 complex values per thread):
//   P1  radix-32 over the high 5 column bits      thread (r, n2c),            c = 4 n1c + n2c
//   P2  radix-4 (low column bits) x radix-8 (high row bits), with the 3 + 7 thread-dependent
//       inter-pass twiddles                       thread (k1c, n2r),          r = 16 n1r + n2r
//   P3  2 x radix-16 over the low row bits         thread (k1r, k1c, k2c_t),   k2c = b | k2c_t << 1
// Forward is decimation in frequency (bit-reversed out), inverse decimation in time (bit-reversed
// in): after forward() register x[b*16 + pp] of a thread holds the UNNORMALISED spectrum at
// (kr, kc) = (p3_kr(pp), p3_kc(b)); inverse() consumes the same arrangement and returns N^2 times
// the inverse transform in the P1 arrangement x[n1c] <-> (r, 4 n1c + n2c).
// tools/fft_decomp_model.py is the NumPy model of the decomposition and proves the exchange
// layouts bank-conflict free.  Replaces jnp.fft.fftn / ifftn (solvers.py:63, :107-114).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "regfft.cuh"

namespace pdeopt {

constexpr int kN = 128;
constexpr int kThreads = 512;

// ---- shared-memory layouts (validated in tools/fft_decomp_model.py) ----------------------
__device__ __forceinline__ int nat_idx(int r, int c) {
  const int c1 = (c >> 1) & 1;
  int pos = (c & 1) | ((c >> 2) << 1) | (c1 << 6);
  pos ^= ((r & 3) << 1) ^ (c1 << 3);
  return r * kN + pos;
}
__device__ __forceinline__ int ex_idx(int k1c, int rho, int q) {
  return ((k1c * 32 + (rho >> 2)) * 16) + ((((rho & 3) ^ ((k1c >> 2) & 3))) << 2) + (q ^ (k1c & 3));
}

// ---- FFT passes ---------------------------------------------------------------------------
// P1: thread (r, n2c) owns c = 4*n1c + n2c, n1c = 0..31.
struct P1Map {
  int r, n2c;
  __device__ __forceinline__ P1Map() {
    const int t = threadIdx.x;
    n2c = t & 3;
    r = ((t >> 2) & 3) | ((t >> 4) << 2);
  }
};
// P2: thread (k1c, n2r) owns (q, hi) with rho = 16*hi + n2r.
struct P2Map {
  int k1c, n2r;
  __device__ __forceinline__ P2Map() {
    const int t = threadIdx.x;
    k1c = (t & 3) | (((t >> 4) & 7) << 2);
    n2r = ((t >> 2) & 3) | (((t >> 7) & 3) << 2);
  }
};
// P3: thread (k1r, k1c, k2c_t) owns n2r = 0..15 and k2c = b | (k2c_t << 1), b = 0,1.
struct P3Map {
  int k1r, k1c, k2ct;
  __device__ __forceinline__ P3Map() {
    const int t = threadIdx.x;
    k1c = t & 31;
    k2ct = (t >> 5) & 1;
    k1r = t >> 6;
  }
};

// Shared-memory accesses of the FFT passes: a per-thread base register XORed with a compile-time
// constant (one LOP3) plus a compile-time immediate offset, so no per-element index arithmetic.
// `a` is a shared-window BYTE address; the buffer is 1024-byte aligned so XORs of low bits commute
// with the base.
template <int OFF>
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ void sts64(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "f"(v.x), "f"(v.y) : "memory");
}

// natural layout, P1 map: element(r, c = 4n + n2c) = (row_base + tbn) ^ (n << 1)   [elements]
__device__ __forceinline__ uint32_t p1_nat_base(uint32_t wbase, const P1Map& m) {
  const int c1 = m.n2c >> 1;
  const int tbn = (m.n2c & 1) | (c1 << 6) | ((m.r & 3) << 1) | (c1 << 3);
  return wbase + (uint32_t)(m.r * kN + tbn) * 8u;
}
__device__ __forceinline__ void p1_gather_nat(uint32_t nb, float2 (&x)[32]) {
  static_for<0, 8>([&](auto lc) {
    constexpr int lo = decltype(lc)::value;
    const uint32_t a = nb ^ (uint32_t)(lo << 4);  // (n & 7) << 1 elements = << 4 bytes
    static_for<0, 4>([&](auto hc) {
      constexpr int hi = decltype(hc)::value;
      x[hi * 8 + lo] = lds64<hi * 8 * 2 * 8>(a);  // (n >> 3) << 4 elements
    });
  });
}
__device__ __forceinline__ void p1_scatter_nat(uint32_t nb, const float2 (&x)[32]) {
  static_for<0, 8>([&](auto lc) {
    constexpr int lo = decltype(lc)::value;
    const uint32_t a = nb ^ (uint32_t)(lo << 4);
    static_for<0, 4>([&](auto hc) {
      constexpr int hi = decltype(hc)::value;
      sts64<hi * 8 * 2 * 8>(a, x[hi * 8 + lo]);
    });
  });
}
// exchange layout, P1 map: element(k1c, r, n2c) = (tb1 ^ Clo(k1c)) + 512 k1c, Clo = k1c & 15
__device__ __forceinline__ uint32_t p1_ex_base(uint32_t wbase, const P1Map& m) {
  return wbase + (uint32_t)((m.r >> 2) * 16 + ((m.r & 3) << 2) + m.n2c) * 8u;
}

struct Fft128 {
  P1Map m1;
  P2Map m2;
  P3Map m3;
  uint32_t nb, e1, tb2, tb3;
  const float2* tw;  // 128-entry table of w_128^e in shared memory

  __device__ __forceinline__ Fft128(uint32_t wbase, const float2* tw_) : tw(tw_) {
    nb = p1_nat_base(wbase, m1);
    e1 = p1_ex_base(wbase, m1);
    // P2: element(k1c, 16 hi + n2r, q) = tb2 + (q ^ (k1c & 3)) + 64 hi
    tb2 = wbase + (uint32_t)(m2.k1c * 512 + (m2.n2r >> 2) * 16 + (((m2.n2r & 3) ^ ((m2.k1c >> 2) & 3)) << 2)) * 8u;
    // P3: element(k1c, 16 k1r + n, k2c = b | k2ct << 1) = tb3 ^ (((n & 3) << 2) | b) + 16 (n >> 2)
    tb3 = wbase + (uint32_t)(m3.k1c * 512 + m3.k1r * 64 + ((((m3.k1c >> 2) & 3)) << 2) + (((m3.k2ct << 1) ^ (m3.k1c & 2))) +
                             (m3.k1c & 1)) * 8u;
  }
  // spectral index held in register x[b*16 + pp] after forward()
  __device__ __forceinline__ int p3_kc(int b) const { return m3.k1c + 32 * (b | (m3.k2ct << 1)); }
  __device__ __forceinline__ int p3_kr(int pp) const { return m3.k1r + 8 * brev<4>(pp); }
  // spatial index held in register x[n] in the P1 arrangement
  __device__ __forceinline__ int p1_row() const { return m1.r; }
  __device__ __forceinline__ int p1_col(int n) const { return 4 * n + m1.n2c; }

  // x: P1 arrangement (natural n1c order).  The exchange buffer must be free (all earlier
  // readers past a barrier), except that a thread may still "own" its own P1 addresses.
  __device__ __forceinline__ void forward(float2 (&x)[32]) const {
    // ---- P1 forward: 32-point DFT over n1c ----
    Dif<32, 1, false>::run(x);
    static_for<0, 16>([&](auto cc) {
      constexpr int clo = decltype(cc)::value;  // k1c & 15
      const uint32_t a = e1 ^ (uint32_t)(clo * 8);
      sts64<clo * 512 * 8>(a, x[brev<5>(clo)]);
      sts64<(clo + 16) * 512 * 8>(a, x[brev<5>(clo + 16)]);
    });
    __syncthreads();
    // ---- P2 forward: twiddle, 4-point DFT over n2c, 8-point DFT over n1r, twiddle ----
    {
      static_for<0, 4>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const uint32_t a = tb2 + (uint32_t)((q ^ (m2.k1c & 3)) * 8);
        static_for<0, 8>([&](auto hc) {
          constexpr int hi = decltype(hc)::value;
          x[q * 8 + hi] = lds64<hi * 64 * 8>(a);
        });
      });
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        const float2 w = tw[(q * m2.k1c) & 127];
#pragma unroll
        for (int hi = 0; hi < 8; ++hi) x[q * 8 + hi] = cmul(x[q * 8 + hi], w);
      }
      static_for<0, 8>([&](auto hc) { Dif<4, 8, false>::run(x + decltype(hc)::value); });
      static_for<0, 4>([&](auto qc) { Dif<8, 1, false>::run(x + 8 * decltype(qc)::value); });
      // position (pq, pr) holds k2c = brev2(pq), k1r = brev3(pr)
      static_for<1, 8>([&](auto rc) {
        constexpr int pr = decltype(rc)::value;
        const float2 w = tw[(brev<3>(pr) * m2.n2r) & 127];
        static_for<0, 4>([&](auto qc) {
          constexpr int pq = decltype(qc)::value;
          x[pq * 8 + pr] = cmul(x[pq * 8 + pr], w);
        });
      });
      static_for<0, 4>([&](auto qc) {
        constexpr int pq = decltype(qc)::value;
        const uint32_t a = tb2 + (uint32_t)((brev<2>(pq) ^ (m2.k1c & 3)) * 8);
        static_for<0, 8>([&](auto rc) {
          constexpr int pr = decltype(rc)::value;
          sts64<brev<3>(pr) * 64 * 8>(a, x[pq * 8 + pr]);
        });
      });
    }
    __syncthreads();
    // ---- P3: 16-point DFT over n2r ----
    static_for<0, 2>([&](auto bc) {
        constexpr int b = decltype(bc)::value;
        static_for<0, 4>([&](auto mc) {
          constexpr int mm = decltype(mc)::value;
          const uint32_t a = tb3 ^ (uint32_t)(((mm << 2) | b) * 8);
          static_for<0, 4>([&](auto hc) {
            constexpr int nh = decltype(hc)::value;
            x[b * 16 + nh * 4 + mm] = lds64<nh * 16 * 8>(a);
          });
        });
      });
    Dif<16, 1, false>::run(x);
    Dif<16, 1, false>::run(x + 16);
  }

  // x: spectrum in the P3 arrangement; returns N^2 * ifft in the P1 arrangement.
  __device__ __forceinline__ void inverse(float2 (&x)[32]) const {
    Dit<16, 1, true>::run(x);
      Dit<16, 1, true>::run(x + 16);
      static_for<0, 2>([&](auto bc) {
        constexpr int b = decltype(bc)::value;
        static_for<0, 4>([&](auto mc) {
          constexpr int mm = decltype(mc)::value;
          const uint32_t a = tb3 ^ (uint32_t)(((mm << 2) | b) * 8);
          static_for<0, 4>([&](auto hc) {
            constexpr int nh = decltype(hc)::value;
            sts64<nh * 16 * 8>(a, x[b * 16 + nh * 4 + mm]);
          });
        });
      });
    __syncthreads();
    // ---- P2 inverse ----
    {
      static_for<0, 4>([&](auto qc) {
        constexpr int pq = decltype(qc)::value;
        const uint32_t a = tb2 + (uint32_t)((brev<2>(pq) ^ (m2.k1c & 3)) * 8);
        static_for<0, 8>([&](auto rc) {
          constexpr int pr = decltype(rc)::value;
          x[pq * 8 + pr] = lds64<brev<3>(pr) * 64 * 8>(a);
        });
      });
      static_for<1, 8>([&](auto rc) {
        constexpr int pr = decltype(rc)::value;
        const float2 w = tw[(brev<3>(pr) * m2.n2r) & 127];
        static_for<0, 4>([&](auto qc) {
          constexpr int pq = decltype(qc)::value;
          x[pq * 8 + pr] = cmulc(x[pq * 8 + pr], w);
        });
      });
      static_for<0, 4>([&](auto qc) { Dit<8, 1, true>::run(x + 8 * decltype(qc)::value); });
      static_for<0, 8>([&](auto hc) { Dit<4, 8, true>::run(x + decltype(hc)::value); });
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        const float2 w = tw[(q * m2.k1c) & 127];
#pragma unroll
        for (int hi = 0; hi < 8; ++hi) x[q * 8 + hi] = cmulc(x[q * 8 + hi], w);
      }
      static_for<0, 4>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const uint32_t a = tb2 + (uint32_t)((q ^ (m2.k1c & 3)) * 8);
        static_for<0, 8>([&](auto hc) {
          constexpr int hi = decltype(hc)::value;
          sts64<hi * 64 * 8>(a, x[q * 8 + hi]);
        });
      });
    }
    __syncthreads();
    // ---- P1 inverse ----
    static_for<0, 16>([&](auto cc) {
      constexpr int clo = decltype(cc)::value;
      const uint32_t a = e1 ^ (uint32_t)(clo * 8);
      x[brev<5>(clo)] = lds64<clo * 512 * 8>(a);
      x[brev<5>(clo + 16)] = lds64<(clo + 16) * 512 * 8>(a);
    });
    Dit<32, 1, true>::run(x);
  }
};

}  // namespace pdeopt
