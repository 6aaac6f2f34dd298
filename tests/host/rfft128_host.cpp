// Host harness for pde_opt_b200/csrc/rfft128.cuh (built by tests/test_rfft128_host.py with g++):
// the 256 threads of the CTA are emulated one barrier phase at a time on a byte buffer that stands
// in for shared memory, so the thread maps, exchange layouts, twiddles and the closed-form filter
// are checked on the CPU exactly as the kernel executes them.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../pde_opt_b200/csrc/rfft128.cuh"
using namespace pdeopt;
using namespace pdeopt::rf;

extern "C" {
// f, g: [128][128] float; tab: [65][65] folded A*sigma; returns g = Re ifft2(fft2(f) / (1 + dt tab)).
// round_trip != 0: skip the filter arithmetic check and also return the scatter/gather round trip in rt.
void rfft128_filter(const float* f, const float* tab, float dt, float* g, float* rt) {
  std::vector<unsigned char> smem(kWBytes + kTRows * kTCols * 16);
  g_emul = smem.data();
  float2 twb[8 * 16], tw64[32], sc[32];
  for (int n2r = 0; n2r < 8; ++n2r)
    for (int k1r = 0; k1r < 16; ++k1r) {
      const double a = -2.0 * M_PI * double(n2r * k1r) / 128.0;
      twb[k1r * 8 + n2r] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
  for (int k = 0; k < 32; ++k) {
    const double a = -2.0 * M_PI * double(k) / 64.0;
    tw64[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    const double b = 2.0 * M_PI * double(k) / 128.0;
    sc[k] = make_float2((float)std::cos(b), (float)std::sin(b));
  }
  float2* W = reinterpret_cast<float2*>(smem.data());
  float4* T = reinterpret_cast<float4*>(smem.data() + kWBytes);
  for (int r = 0; r < kRows; ++r)
    for (int m = 0; m < kH; ++m) W[nat_slot(r, m)] = make_float2(f[r * kCols + 2 * m], f[r * kCols + 2 * m + 1]);
  for (int fr = 0; fr < kTRows; ++fr)
    for (int c = 0; c < kTCols; ++c) T[fr * kTCols + c] = filter_entry(tab, sc, fr, c, dt);
  std::vector<float2> xs(kThreadsR * 32);
  auto X = [&](int t) -> float2(&)[32] { return *reinterpret_cast<float2(*)[32]>(&xs[t * 32]); };
  std::vector<RFft> F;
  for (int t = 0; t < kThreadsR; ++t) F.emplace_back(0u, kWBytes, t);
  for (int t = 0; t < kThreadsR; ++t) gather_nat<true>(F[t], X(t));
  for (int t = 0; t < kThreadsR; ++t) passA_fwd(F[t], X(t));
  for (int t = 0; t < kThreadsR; ++t) passB_fwd(F[t], twb, tw64, X(t));
  for (int t = 0; t < kThreadsR; ++t) passC_filter(F[t], X(t));
  for (int t = 0; t < kThreadsR; ++t) passB_inv(F[t], twb, tw64, X(t));
  for (int t = 0; t < kThreadsR; ++t) passA_inv(F[t], X(t));
  for (int t = 0; t < kThreadsR; ++t) {
    const int m0 = t & 1, r = t >> 1;
    for (int n = 0; n < 32; ++n) {
      g[r * kCols + 2 * (2 * n + m0)] = X(t)[n].x;
      g[r * kCols + 2 * (2 * n + m0) + 1] = X(t)[n].y;
    }
  }
  if (rt) {
    for (int t = 0; t < kThreadsR; ++t) scatter_nat(F[t], X(t));
    for (int r = 0; r < kRows; ++r)
      for (int m = 0; m < kH; ++m) {
        rt[r * kCols + 2 * m] = W[nat_slot(r, m)].x;
        rt[r * kCols + 2 * m + 1] = W[nat_slot(r, m)].y;
      }
  }
  g_emul = nullptr;
}
}
