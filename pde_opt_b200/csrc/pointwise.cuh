// Pointwise closure families for mu_h(c), D(c), R(c).
//
// The reference takes arbitrary Python callables for `mu`, `D`, `R`
// (equations/cahn_hilliard.py:50-53, allen_cahn.py:47-50); callables cannot cross a C ABI,
// so the families that the reference's tests, notebooks and docs actually use are
// enumerated here and evaluated inside the fused kernels.
#pragma once
#include <cuda_runtime.h>

#include "regfft.cuh"

namespace pdeopt {

enum : int { MU_DOUBLE_WELL = 0, MU_LOG = 1, MU_LEGENDRE = 2, MU_LEGENDRE_LOGPRIOR = 3, MU_RUNTIME = -1 };
enum : int { MOB_CONST = 0, MOB_DEGENERATE = 1, MOB_ONE_PLUS_SQ = 2, MOB_LEGENDRE_EXP = 3, MOB_RUNTIME = -1 };

struct PointwiseParams {
  int mu_family, mu_ncoef;
  float mu_coef[16];
  int mob_family, mob_ncoef;
  float mob_coef[16];
};

// Legendre recurrence constants: P_n = a_n x P_{n-1} - b_n P_{n-2} with a_n = (2n-1)/n, b_n = (n-1)/n.  The reference
// divides by n (functions/legendre.py:27-31); multiplying by the rounded reciprocal constants differs from that by
// an ulp or two per term and saves an IEEE division (about twenty instructions) per term and grid point.
template <int N>
struct LegC {
  static constexpr float a = float(double(2 * N - 1) / double(N));
  static constexpr float b = float(double(N - 1) / double(N));
};

// functions/legendre.py:19-34: three-term recurrence on x in [-1,1] (at most 16 coefficients).
__device__ __forceinline__ float legendre_eval(const float* __restrict__ coef, int ncoef, float x) {
  float result = coef[0];
  if (ncoef > 1) result = fmaf(coef[1], x, result);
  float p_prev = 1.0f, p_curr = x;
  static_for<2, 16>([&](auto nc) {
    constexpr int n = decltype(nc)::value;
    if (n < ncoef) {
      const float p_next = LegC<n>::a * x * p_curr - LegC<n>::b * p_prev;
      result = fmaf(coef[n], p_next, result);
      p_prev = p_curr;
      p_curr = p_next;
    }
  });
  return result;
}

// sum_n a[n] P_n(x) and sum_n b[n] P_n(x) over ONE pass of the recurrence (mu and the mobility of the same point)
__device__ __forceinline__ void legendre_eval2(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, float x,
                                               float& ra, float& rb) {
  ra = na > 0 ? a[0] : 0.0f;
  rb = nb > 0 ? b[0] : 0.0f;
  if (na > 1) ra = fmaf(a[1], x, ra);
  if (nb > 1) rb = fmaf(b[1], x, rb);
  const int nmax = na > nb ? na : nb;
  float p_prev = 1.0f, p_curr = x;
  static_for<2, 16>([&](auto nc) {
    constexpr int n = decltype(nc)::value;
    if (n < nmax) {
      const float p_next = LegC<n>::a * x * p_curr - LegC<n>::b * p_prev;
      if (n < na) ra = fmaf(a[n], p_next, ra);
      if (n < nb) rb = fmaf(b[n], p_next, rb);
      p_prev = p_curr;
      p_curr = p_next;
    }
  });
}

// log(c/(1-c)): one MUFU.RCP-based division and one MUFU.LG2.
__device__ __forceinline__ float logit(float c) { return __logf(__fdividef(c, 1.0f - c)); }

template <int MU>
__device__ __forceinline__ float mu_h(float c, const PointwiseParams& pw, float w_off) {
  const int fam = (MU == MU_RUNTIME) ? pw.mu_family : MU;
  if (fam == MU_DOUBLE_WELL) {
    return c * c * c - c;  // tests/test_solvers.py:36
  } else if (fam == MU_LOG) {
    return logit(c) + (pw.mu_coef[0] + w_off) * (1.0f - 2.0f * c);  // optimize_nn_script.py:33
  } else if (fam == MU_LEGENDRE) {
    return legendre_eval(pw.mu_coef, pw.mu_ncoef, 2.0f * c - 1.0f);  // legendre.py:68-73
  } else {
    return legendre_eval(pw.mu_coef, pw.mu_ncoef, 2.0f * c - 1.0f) + logit(c);
  }
}

template <int MOB>
__device__ __forceinline__ float mob(float c, const PointwiseParams& pw) {
  const int fam = (MOB == MOB_RUNTIME) ? pw.mob_family : MOB;
  if (fam == MOB_CONST) {
    return pw.mob_coef[0];
  } else if (fam == MOB_DEGENERATE) {
    return (1.0f - c) * c;
  } else if (fam == MOB_ONE_PLUS_SQ) {
    return 1.0f + c * c;  // test_rhs_convergence.py:22,55
  } else {
    return __expf(legendre_eval(pw.mob_coef, pw.mob_ncoef, 2.0f * c - 1.0f));  // legendre.py:48-53
  }
}

// ---- packed (two environments per register pair) evaluation -------------------------------
__device__ __forceinline__ float lg2_fast(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// mu_h(c) and mobility D(c) for the (a, b) pair held in one float2.  `w` is the per-env value of
// the first mu coefficient (family coefficient + control offset).
template <int MU, int MOB>
__device__ __forceinline__ void mu_mob_pair(float2 c, const PointwiseParams& pw, float2 w_off, float2& mu, float2& D) {
  if constexpr (MU == MU_LOG && (MOB == MOB_DEGENERATE || MOB == MOB_CONST)) {
    const float2 s = sub2(splat2(1.0f), c);                                 // 1 - c
    const float2 l = sub2(make_float2(lg2_fast(c.x), lg2_fast(c.y)),        // log2(c) - log2(1-c)
                          make_float2(lg2_fast(s.x), lg2_fast(s.y)));
    const float2 t = fma2(c, splat2(-2.0f), splat2(1.0f));                  // 1 - 2c
    const float2 w = add2(splat2(pw.mu_coef[0]), w_off);
    mu = fma2(l, splat2(0.69314718055994531f), mul2(w, t));
    D = (MOB == MOB_DEGENERATE) ? mul2(s, c) : splat2(pw.mob_coef[0]);
  } else if constexpr (MU == MU_DOUBLE_WELL && (MOB == MOB_CONST || MOB == MOB_ONE_PLUS_SQ)) {
    const float2 c2 = mul2(c, c);
    mu = sub2(mul2(c2, c), c);  // c^3 - c
    D = (MOB == MOB_CONST) ? splat2(pw.mob_coef[0]) : add2(splat2(1.0f), c2);
  } else {
    const int mf = (MU == MU_RUNTIME) ? pw.mu_family : MU, df = (MOB == MOB_RUNTIME) ? pw.mob_family : MOB;
    if ((mf == MU_LEGENDRE || mf == MU_LEGENDRE_LOGPRIOR) && df == MOB_LEGENDRE_EXP) {
      // the training closures (legendre.py:37-74): mu and D of a point share one pass over the Legendre basis
      float m0, d0, m1, d1;
      legendre_eval2(pw.mu_coef, pw.mu_ncoef, pw.mob_coef, pw.mob_ncoef, 2.0f * c.x - 1.0f, m0, d0);
      legendre_eval2(pw.mu_coef, pw.mu_ncoef, pw.mob_coef, pw.mob_ncoef, 2.0f * c.y - 1.0f, m1, d1);
      if (mf == MU_LEGENDRE_LOGPRIOR) {
        m0 += logit(c.x);
        m1 += logit(c.y);
      }
      mu = make_float2(m0, m1);
      D = make_float2(__expf(d0), __expf(d1));
    } else {
      mu = make_float2(mu_h<MU>(c.x, pw, w_off.x), mu_h<MU>(c.y, pw, w_off.y));
      D = make_float2(mob<MOB>(c.x, pw), mob<MOB>(c.y, pw));
    }
  }
}

}  // namespace pdeopt
