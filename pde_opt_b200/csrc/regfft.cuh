// In-register power-of-two FFTs on float2 arrays with compile-time twiddles.
//
// Building block of the fused in-SMEM spectral steppers: every pass of the 2-D transform
// is a set of small DFTs held entirely in one thread's registers (fully unrolled, all
// indices compile-time so the array lives in registers).  Forward transforms are
// decimation-in-frequency (natural order in, bit-reversed order out); inverse transforms
// are decimation-in-time (bit-reversed in, natural out), so a forward/inverse pair needs no
// reordering and a spectral multiplier is simply applied in bit-reversed positions.
//
// Replaces jnp.fft.fftn / ifftn as called from the reference's
// pde_opt/numerics/solvers.py:63 and :107-114 (XLA library FFT there).
//
// The header is host-compilable (tests/test_regfft_host.py builds it with g++).
#pragma once
#include <cmath>
#include <type_traits>

#if defined(__CUDACC__)
#define PDEOPT_HD __host__ __device__ __forceinline__
#else
#define PDEOPT_HD inline
#ifndef PDEOPT_HOST_FLOAT2
#define PDEOPT_HOST_FLOAT2
struct float2 {
  float x, y;
};
static inline float2 make_float2(float a, float b) {
  float2 r;
  r.x = a;
  r.y = b;
  return r;
}
#endif
#endif

namespace pdeopt {

// ---- compile-time sin/cos (Taylor on [-pi, pi], double precision) ----------------------
constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double cx_sin(double x) {
  double x2 = x * x, term = x, sum = x;
  for (int i = 1; i < 16; ++i) {
    term *= -x2 / double((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
constexpr double cx_cos(double x) {
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int i = 1; i < 16; ++i) {
    term *= -x2 / double((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}

// w_N^J = exp(-2 pi i J / N) (forward) ; conjugate for the inverse.
template <int N, int J>
struct Tw {
  static constexpr float re = float(cx_cos(2.0 * kPi * double(J) / double(N)));
  static constexpr float im = float(-cx_sin(2.0 * kPi * double(J) / double(N)));
};

template <int LOG2N>
constexpr int brev(int k) {
  int r = 0;
  for (int b = 0; b < LOG2N; ++b) r |= ((k >> b) & 1) << (LOG2N - 1 - b);
  return r;
}

constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// Packed two-lane FP32 arithmetic (sm_100 add/sub/mul/fma .f32x2: one issue slot, both lanes).
// A complex add is exactly one packed add; on the host the scalar form is used.
#if defined(__CUDA_ARCH__) && !defined(PDEOPT_NO_F32X2)
__device__ __forceinline__ unsigned long long pk2(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 unpk2(unsigned long long u) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(u));
  return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return unpk2(r);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return unpk2(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return unpk2(r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
  return unpk2(r);
}
#else
PDEOPT_HD float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PDEOPT_HD float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
PDEOPT_HD float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
PDEOPT_HD float2 fma2(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
PDEOPT_HD float2 cadd(float2 a, float2 b) { return add2(a, b); }
PDEOPT_HD float2 csub(float2 a, float2 b) { return sub2(a, b); }
// Complex products in packed form.  ptxas folds the lane swap (b.y, b.x) and the sign pattern into
// operand modifiers of FFMA2 / FMUL2 (.LO_HI, .NP) and broadcasts the scalar factor (.F32 operand), so
// a complex multiply is two packed instructions and a complex multiply-add  a + w b  two FFMA2 — no
// scalar FP32 instruction.  That matters on sm_100: a scalar FP32 instruction issued between packed
// ones occupies one 16-lane half of the FMA pipe for two cycles while the other half idles
// (tools/issue_bench.cu: 8 FADD2 + 8 FFMA per warp take 124.6 cycles per scheduler, not 96).
// Same products and the same rounding order as the scalar forms.
#if defined(__CUDA_ARCH__) && !defined(PDEOPT_NO_F32X2)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  const float2 t = mul2(make_float2(-w.y, w.y), make_float2(a.y, a.x));
  return fma2(make_float2(w.x, w.x), a, t);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 w) {  // a * conj(w)
  const float2 t = mul2(make_float2(w.y, -w.y), make_float2(a.y, a.x));
  return fma2(make_float2(w.x, w.x), a, t);
}
// a + w b
__device__ __forceinline__ float2 cmac(float2 a, float2 b, float wr, float wi) {
  const float2 t = fma2(make_float2(-wi, wi), make_float2(b.y, b.x), a);
  return fma2(make_float2(wr, wr), b, t);
}
// 2a - p
__device__ __forceinline__ float2 twice_minus(float2 a, float2 p) {
  return fma2(make_float2(2.0f, 2.0f), a, make_float2(-p.x, -p.y));
}
#else
PDEOPT_HD float2 cmul(float2 a, float2 w) {
  return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}
PDEOPT_HD float2 cmulc(float2 a, float2 w) {  // a * conj(w)
  return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}
PDEOPT_HD float2 cmac(float2 a, float2 b, float wr, float wi) {
  return make_float2(fmaf(wr, b.x, fmaf(-wi, b.y, a.x)), fmaf(wr, b.y, fmaf(wi, b.x, a.y)));
}
PDEOPT_HD float2 twice_minus(float2 a, float2 p) { return make_float2(fmaf(2.0f, a.x, -p.x), fmaf(2.0f, a.y, -p.y)); }
#endif

// (a) * w_N^J, with the trivial and 8th-root cases specialised at compile time.
template <int N, int J, bool INV>
PDEOPT_HD float2 mul_tw(float2 a) {
  constexpr int Jm = ((J % N) + N) % N;
  if constexpr (Jm == 0) {
    return a;
  } else if constexpr (Jm * 4 == N) {  // -i (fwd) / +i (inv)
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
  } else if constexpr (Jm * 2 == N) {
    return make_float2(-a.x, -a.y);
  } else if constexpr (Jm * 4 == 3 * N) {
    return INV ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
  } else if constexpr (Jm * 8 == N) {  // (1 - i)/sqrt2 fwd
    constexpr float s = 0.70710678118654752440f;
    return INV ? make_float2((a.x - a.y) * s, (a.x + a.y) * s) : make_float2((a.x + a.y) * s, (a.y - a.x) * s);
  } else if constexpr (Jm * 8 == 3 * N) {  // (-1 - i)/sqrt2 fwd
    constexpr float s = 0.70710678118654752440f;
    return INV ? make_float2(-(a.x + a.y) * s, (a.x - a.y) * s) : make_float2((a.y - a.x) * s, -(a.x + a.y) * s);
  } else {
    constexpr float wr = Tw<N, Jm>::re;
    constexpr float wi = INV ? -Tw<N, Jm>::im : Tw<N, Jm>::im;
    return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
  }
}

// true when w_N^J is none of the cases mul_tw specialises (1, -1, +-i, 8th roots)
template <int N, int J>
constexpr bool tw_is_general() {
  constexpr int Jm = ((J % N) + N) % N;
  return !(Jm == 0 || Jm * 4 == N || Jm * 2 == N || Jm * 4 == 3 * N || Jm * 8 == N || Jm * 8 == 3 * N || Jm * 8 == 5 * N ||
           Jm * 8 == 7 * N);
}

template <int I, int E, class F>
PDEOPT_HD void static_for(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(static_cast<F&&>(f));
  }
}

// Decimation in frequency, in place on x[0], x[S], ..., x[(N-1)S].
// Natural order in; position p (units of S) holds X[brev(p)] on return.
template <int N, int S, bool INV>
struct Dif {
  static PDEOPT_HD void run(float2* x) {
    if constexpr (N >= 2) {
      static_for<0, N / 2>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        float2 a = x[j * S], b = x[(j + N / 2) * S];
        x[j * S] = cadd(a, b);
        x[(j + N / 2) * S] = mul_tw<N, j, INV>(csub(a, b));
      });
      Dif<N / 2, S, INV>::run(x);
      Dif<N / 2, S, INV>::run(x + (N / 2) * S);
    }
  }
};

// Decimation in time, in place.  Position p holds x[brev(p)] on entry; natural order out.
template <int N, int S, bool INV>
struct Dit {
  static PDEOPT_HD void run(float2* x) {
    if constexpr (N >= 2) {
      Dit<N / 2, S, INV>::run(x);
      Dit<N / 2, S, INV>::run(x + (N / 2) * S);
      static_for<0, N / 2>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (tw_is_general<N, j>()) {
          // a + w b by four FMAs, a - w b = 2a - (a + w b) by two more: six pipe slots per butterfly
          // instead of eight (the FP32 pipe charges an add like an FMA)
          constexpr float wr = Tw<N, j>::re;
          constexpr float wi = INV ? -Tw<N, j>::im : Tw<N, j>::im;
          const float2 a = x[j * S], b = x[(j + N / 2) * S];
          float2 p;
          p.x = fmaf(wr, b.x, fmaf(-wi, b.y, a.x));
          p.y = fmaf(wr, b.y, fmaf(wi, b.x, a.y));
          x[j * S] = p;
          x[(j + N / 2) * S] = make_float2(fmaf(2.0f, a.x, -p.x), fmaf(2.0f, a.y, -p.y));
        } else {
          float2 a = x[j * S], b = mul_tw<N, j, INV>(x[(j + N / 2) * S]);
          x[j * S] = cadd(a, b);
          x[(j + N / 2) * S] = csub(a, b);
        }
      });
    }
  }
};

// Decimation in time with EVERY non-trivial twiddle (the 8th roots included) in the fused form
// p = a + w b (four FMAs), q = 2a - p (two FMAs): six FP32-pipe slots per butterfly instead of eight,
// four for the trivial twiddles 1 and -+i.  Position p holds x[brev(p)] on entry; natural order out.
// Used for the forward AND inverse passes of the one-field-per-CTA kernel (register indices are
// compile-time, so the bit-reversed placement of the inputs costs nothing).
template <int N, int S, bool INV>
struct DitF {
  // butterfly J of the LAST stage (combines the two half-size transforms): x[J], x[J + N/2]
  template <int J>
  static PDEOPT_HD void last(float2* x) {
    if constexpr (J == 0) {
      const float2 a = x[0], b = x[(N / 2) * S];
      x[0] = cadd(a, b);
      x[(N / 2) * S] = csub(a, b);
    } else if constexpr (J * 4 == N) {
      const float2 a = x[J * S], b = mul_tw<N, J, INV>(x[(J + N / 2) * S]);
      x[J * S] = cadd(a, b);
      x[(J + N / 2) * S] = csub(a, b);
    } else {
      constexpr float wr = Tw<N, J>::re;
      constexpr float wi = INV ? -Tw<N, J>::im : Tw<N, J>::im;
      const float2 a = x[J * S], b = x[(J + N / 2) * S];
      const float2 p = cmac(a, b, wr, wi);
      x[J * S] = p;
      x[(J + N / 2) * S] = twice_minus(a, p);
    }
  }
  // everything but the last stage
  static PDEOPT_HD void halves(float2* x) {
    if constexpr (N >= 2) {
      DitF<N / 2, S, INV>::run(x);
      DitF<N / 2, S, INV>::run(x + (N / 2) * S);
    }
  }
  static PDEOPT_HD void run(float2* x) {
    if constexpr (N >= 2) {
      halves(x);
      static_for<0, N / 2>([&](auto jc) { last<decltype(jc)::value>(x); });
    }
  }
};

}  // namespace pdeopt
