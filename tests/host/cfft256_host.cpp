// Host harness for pde_opt_b200/csrc/cfft256.cuh (built by tests/test_cfft256_host.py with g++): a
// cluster of 4 CTAs x 512 threads is emulated one barrier phase at a time on four byte buffers standing
// in for the CTAs' shared memory (a remote store = a write into the peer's buffer), so the thread maps,
// the slab layout, the in-line exchange and the transposed stores are checked on the CPU.
#include <cmath>
#include <vector>

#include "../../pde_opt_b200/csrc/cfft256.cuh"
using namespace pdeopt;
using namespace pdeopt::cf;

extern "C" {
// in/out: [256][256][2] float (re, im), natural order.  mult: [256][256][2] complex multiplier in natural
// frequency order [kr][kc] (or null).  out = ifft2(fft2(in) * mult) * 65536 (unnormalised inverse);
// spec (optional, [256][256][2]): the forward transform as the column pass sees it.
void cfft256_roundtrip(const float* in, const float* mult, float* out, float* spec) {
  std::vector<std::vector<unsigned char>> slab(kCtas, std::vector<unsigned char>(kSlabBytes));
  float2 tw[8 * 32];
  for (int j = 0; j < 8; ++j)
    for (int k1 = 0; k1 < 32; ++k1) {
      const double a = -2.0 * M_PI * double(j * k1) / 256.0;
      tw[j * 32 + k1] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
  std::vector<float2> regs((size_t)kCtas * kThreadsC * 32);
  auto X = [&](int q, int t) -> float2(&)[32] { return *reinterpret_cast<float2(*)[32]>(&regs[((size_t)q * kThreadsC + t) * 32]); };
  auto ctx = [&](int q) { Ctx c; for (int r = 0; r < kCtas; ++r) c.slabs[r] = slab[r].data(); c.rank = q; return c; };
#define FOR_ALL(...) for (int q = 0; q < kCtas; ++q) { Ctx c = ctx(q); (void)c; for (int t = 0; t < kThreadsC; ++t) { const int l = thread_line(t), j = thread_j(t); const LineMap m(c, l, j, kLines * q + l); const float2* tj = tw + 32 * j; (void)m; (void)tj; __VA_ARGS__; } }
  // row slabs: CTA q, line l = row 64 q + l
  FOR_ALL(static_for<0, 32>([&](auto nc) { constexpr int n1 = decltype(nc)::value; const int r = kLines * q + l, col = 8 * n1 + j;
                                             X(q, t)[brev<5>(n1)] = make_float2(in[(r * 256 + col) * 2], in[(r * 256 + col) * 2 + 1]); }))
  // row pass forward, transposed store -> column slabs
  FOR_ALL(line_fwd_a(c, tj, m, X(q, t)))
  FOR_ALL(line_fwd_b(c, m, X(q, t)))
  FOR_ALL(store_transposed_from_freq(c, m, X(q, t)))
  // column pass: forward, multiply, inverse
  FOR_ALL(load_spatial(c, m, X(q, t)))
  FOR_ALL(line_fwd_a(c, tj, m, X(q, t)))
  FOR_ALL(line_fwd_b(c, m, X(q, t)))
  FOR_ALL(for (int i = 0; i < 32; ++i) { const int kr = j + 8 * (i >> 3) + 32 * (i & 7), kc = kLines * q + l;
            if (spec) { spec[(kr * 256 + kc) * 2] = X(q, t)[i].x; spec[(kr * 256 + kc) * 2 + 1] = X(q, t)[i].y; }
            if (mult) X(q, t)[i] = cmul(X(q, t)[i], make_float2(mult[(kr * 256 + kc) * 2], mult[(kr * 256 + kc) * 2 + 1])); })
  FOR_ALL(line_inv_a(c, m, X(q, t)))
  FOR_ALL(line_inv_b(c, tj, m, X(q, t)))
  FOR_ALL(store_transposed_from_spatial(c, m, X(q, t)))
  // row pass inverse
  FOR_ALL(load_freq(c, m, X(q, t)))
  FOR_ALL(line_inv_a(c, m, X(q, t)))
  FOR_ALL(line_inv_b(c, tj, m, X(q, t)))
  FOR_ALL(for (int n1 = 0; n1 < 32; ++n1) { const int r = kLines * q + l, col = 8 * n1 + j;
            out[(r * 256 + col) * 2] = X(q, t)[n1].x; out[(r * 256 + col) * 2 + 1] = X(q, t)[n1].y; })
#undef FOR_ALL
}
}
