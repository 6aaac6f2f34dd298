// Helpers shared by the fused kernels built on Fft128 + TMEM parking (ad128.cuh, fourier128.cuh).
#pragma once
#include "fft128.cuh"
#include "sifs128.cuh"

namespace pdeopt {

__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }

__device__ __forceinline__ void park_all(const Park& pk, const float2 (&x)[32]) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = x[ch * 8 + i];
    pk.store(ch, v);
  }
  pk.fence_store();
}

// Spectral bookkeeping for register chunk CH (registers x[CH*8 .. CH*8+7] after Fft128::forward):
// fn(ic, kr, kc, ft) with ic the compile-time index inside the chunk, (kr, kc) the wavenumber
// indices along axis 0 / axis 1 and ft the index into the folded (even) 65x65 tables.
template <int CH, class Fn>
__device__ __forceinline__ void spec_chunk(const Fft128& F, Fn&& fn) {
  constexpr int b = CH >> 1;
  const int kc = F.p3_kc(b);
  const int fc = kc <= 64 ? kc : 128 - kc;
  static_for<0, 8>([&](auto ic) {
    constexpr int pp = (CH & 1) * 8 + decltype(ic)::value;
    const int kr = F.p3_kr(pp);
    const int fr = kr <= 64 ? kr : 128 - kr;
    fn(ic, kr, kc, fr * kTabDim + fc);
  });
}

}  // namespace pdeopt
