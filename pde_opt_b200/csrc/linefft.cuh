// Batched 1-D line FFT engine (sm_100a): power-of-two lines of 8..512 complex points, arbitrary
// strides, tiles of adjacent lines staged in shared memory, radix-8 register butterflies.
//
// This is the building block of the paths whose field does not fit one SM:
//   * 256x256 complex64 Strang split-step (GPE, BASELINE config 3): rows pass + columns pass with
//     the state resident in the 126 MB L2 between kernels,
//   * 3-D Cahn-Hilliard (CahnHilliard3DPeriodic, cahn_hilliard.py:112-200; BASELINE config 5):
//     z, y, x line passes; the x pass runs after the slab all-to-all when the domain is sharded.
// It replaces jnp.fft.fftn / ifftn (call sites: cahn_hilliard.py:156-157, gross_pitaevskii.py:58-59,
// solvers.py:63, :107-114).
//
// Transforms are decimation in frequency forward / decimation in time inverse, in place, with the
// spectrum left in DIGIT-REVERSED position order along each axis (line_pos_to_freq): every
// spectral operator of these steppers is a pointwise multiplier, so no reordering pass is ever
// needed; multiplier tables are permuted once at plan time instead.
//
// Shared-memory tile layout: S[idx * (T + 1) + line], T adjacent lines per tile.  A warp always
// touches consecutive `line` at a fixed idx, so every butterfly access is bank-conflict free, and
// for strided lines (adjacent lines contiguous in memory) the global accesses are coalesced rows
// of T * 8 bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "regfft.cuh"

namespace pdeopt {

// ---- radix plan: N = R1 * R2 * R3 with R1 = min(N, 8) etc. -----------------------------------
constexpr int lf_r1(int n) { return n >= 8 ? 8 : n; }
constexpr int lf_r2(int n) { return lf_r1(n / lf_r1(n)); }
constexpr int lf_r3(int n) { return n / lf_r1(n) / lf_r2(n); }

// storage position -> frequency index after the forward transform
__host__ __device__ inline int line_pos_to_freq(int n, int p) {
  const int r1 = lf_r1(n), r2 = lf_r2(n);
  const int s1 = n / r1, s2 = s1 / r2;
  const int k1 = p / s1, rem = p % s1;
  const int k2 = rem / s2, k3 = rem % s2;
  return k1 + r1 * k2 + r1 * r2 * k3;
}

template <int N>
struct LineTile {
  static constexpr int T = (N >= 512) ? 16 : 32;  // lines per tile
  static constexpr int LP = T + 1;
  static constexpr int kThreads = 256;
  static constexpr size_t smem_bytes = sizeof(float2) * (size_t)(N * LP + N);
};

// One radix-R stage with span S on a tile.  Forward: DFT over m then twiddle w_S^{jk}; inverse:
// conjugate twiddle then inverse DFT (exact mirror), see the derivation in DESIGN.md.
template <int N, int R, int S, bool INV>
__device__ __forceinline__ void lf_stage(float2* __restrict__ Sm, const float2* __restrict__ tw) {
  if constexpr (R > 1) {
    constexpr int T = LineTile<N>::T, LP = LineTile<N>::LP, NT = LineTile<N>::kThreads;
    constexpr int SUB = S / R;  // distance between the R inputs of one butterfly
    constexpr int LOG2R = ilog2(R);
    for (int w = threadIdx.x; w < (N / R) * T; w += NT) {
      const int line = w % T, u = w / T;
      const int blk = u / SUB, j = u % SUB;
      float2* base = Sm + (blk * S + j) * LP + line;
      float2 x[R];
      if constexpr (!INV) {
#pragma unroll
        for (int m = 0; m < R; ++m) x[m] = base[m * SUB * LP];
        Dif<R, 1, false>::run(x);
        static_for<0, R>([&](auto pc) {
          constexpr int p = decltype(pc)::value;
          constexpr int k = brev<LOG2R>(p);
          float2 v = x[p];
          if constexpr (k != 0 && SUB > 1) v = cmul(v, tw[(j * k * (N / S)) & (N - 1)]);
          base[k * SUB * LP] = v;
        });
      } else {
        static_for<0, R>([&](auto pc) {
          constexpr int p = decltype(pc)::value;
          constexpr int k = brev<LOG2R>(p);
          float2 v = base[k * SUB * LP];
          if constexpr (k != 0 && SUB > 1) v = cmulc(v, tw[(j * k * (N / S)) & (N - 1)]);
          x[p] = v;
        });
        Dit<R, 1, true>::run(x);
#pragma unroll
        for (int m = 0; m < R; ++m) base[m * SUB * LP] = x[m];
      }
    }
    __syncthreads();
  }
}

template <int N, bool INV>
__device__ __forceinline__ void lf_transform(float2* Sm, const float2* tw) {
  constexpr int R1 = lf_r1(N), R2 = lf_r2(N), R3 = lf_r3(N);
  if constexpr (!INV) {
    lf_stage<N, R1, N, false>(Sm, tw);
    lf_stage<N, R2, N / R1, false>(Sm, tw);
    lf_stage<N, R3, N / R1 / R2, false>(Sm, tw);
  } else {
    lf_stage<N, R3, N / R1 / R2, true>(Sm, tw);
    lf_stage<N, R2, N / R1, true>(Sm, tw);
    lf_stage<N, R1, N, true>(Sm, tw);
  }
}

// Line addressing (all strides in elements of the addressed array):
//   offset(line, idx) = (line / n_inner) * outer + (line % n_inner) * inner
//                     + (idx / chunk) * hi + (idx % chunk) * lo
// `chunk` < N expresses the packed layout of the slab all-to-all.
struct LineGeom {
  long long n_lines;
  long long n_inner, outer, inner;
  int chunk;
  long long hi, lo;
  __device__ __forceinline__ long long off(long long line, int idx) const {
    return (line / n_inner) * outer + (line % n_inner) * inner + (long long)(idx / chunk) * hi + (long long)(idx % chunk) * lo;
  }
};

enum : int { LF_FWD = 0, LF_INV = 1, LF_FWD_MUL_INV = 2 };

// Loader: float2 load(long long line, int idx) ; Mid: float2 apply(float2 v, long long line, int pos) ;
// Storer: void store(long long line, int idx, float2 v).
// CONTIG: elements of a line are adjacent in memory (lo == 1): a warp reads 32 consecutive idx of
// one line; otherwise adjacent lines are adjacent in memory: a warp reads T lines at one idx.
template <int N, int MODE, bool CONTIG, class Loader, class Mid, class Storer>
__global__ void __launch_bounds__(LineTile<N>::kThreads) linefft_kernel(long long n_lines, Loader ld, Mid mid, Storer st) {
  extern __shared__ __align__(16) unsigned char lf_smem[];
  constexpr int T = LineTile<N>::T, LP = LineTile<N>::LP, NT = LineTile<N>::kThreads;
  float2* Sm = reinterpret_cast<float2*>(lf_smem);
  float2* tw = Sm + N * LP;
  for (int i = threadIdx.x; i < N; i += NT) {
    float s, c;
    sincospif(-2.0f * float(i) / float(N), &s, &c);
    tw[i] = make_float2(c, s);
  }
  for (long long tile = blockIdx.x; tile * T < n_lines; tile += gridDim.x) {
    const long long l0 = tile * T;
    // ---- load ----
    if constexpr (CONTIG) {
      for (int w = threadIdx.x; w < N * T; w += NT) {
        const int idx = w % N, line = w / N;
        if (l0 + line < n_lines) Sm[idx * LP + line] = ld.load(l0 + line, idx);
      }
    } else {
      for (int w = threadIdx.x; w < N * T; w += NT) {
        const int line = w % T, idx = w / T;
        if (l0 + line < n_lines) Sm[idx * LP + line] = ld.load(l0 + line, idx);
      }
    }
    __syncthreads();
    if constexpr (MODE == LF_FWD || MODE == LF_FWD_MUL_INV) lf_transform<N, false>(Sm, tw);
    if constexpr (MODE == LF_FWD_MUL_INV) {
      for (int w = threadIdx.x; w < N * T; w += NT) {
        const int line = w % T, pos = w / T;
        if (l0 + line < n_lines) Sm[pos * LP + line] = mid.apply(Sm[pos * LP + line], l0 + line, pos);
      }
      __syncthreads();
    }
    if constexpr (MODE == LF_INV || MODE == LF_FWD_MUL_INV) lf_transform<N, true>(Sm, tw);
    // ---- store ----
    if constexpr (CONTIG) {
      for (int w = threadIdx.x; w < N * T; w += NT) {
        const int idx = w % N, line = w / N;
        if (l0 + line < n_lines) st.store(l0 + line, idx, Sm[idx * LP + line]);
      }
    } else {
      for (int w = threadIdx.x; w < N * T; w += NT) {
        const int line = w % T, idx = w / T;
        if (l0 + line < n_lines) st.store(l0 + line, idx, Sm[idx * LP + line]);
      }
    }
    st.flush(l0);  // optional per-tile reduction hook (uniform; may contain barriers)
    __syncthreads();
  }
}

// ---- stock functors ---------------------------------------------------------------------------
struct LfLoadC {  // complex array
  const float2* p;
  LineGeom g;
  __device__ __forceinline__ float2 load(long long line, int idx) const { return p[g.off(line, idx)]; }
};
struct LfStoreC {
  float2* p;
  LineGeom g;
  __device__ __forceinline__ void store(long long line, int idx, float2 v) const { p[g.off(line, idx)] = v; }
  __device__ __forceinline__ void flush(long long) {}
};
struct LfStoreCScaled {
  float2* p;
  LineGeom g;
  float scale;
  __device__ __forceinline__ void store(long long line, int idx, float2 v) const {
    p[g.off(line, idx)] = make_float2(v.x * scale, v.y * scale);
  }
  __device__ __forceinline__ void flush(long long) {}
};
struct LfMidNone {
  __device__ __forceinline__ float2 apply(float2 v, long long, int) const { return v; }
};

template <int N, int MODE, bool CONTIG, class Loader, class Mid, class Storer>
cudaError_t lf_launch(long long n_lines, Loader ld, Mid mid, Storer st, cudaStream_t stream) {
  auto kern = linefft_kernel<N, MODE, CONTIG, Loader, Mid, Storer>;
  constexpr size_t smem = LineTile<N>::smem_bytes;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const long long tiles = (n_lines + LineTile<N>::T - 1) / LineTile<N>::T;
  const int per_sm = (int)((227 * 1024) / (smem + 1024));
  long long grid = tiles;
  const long long cap = 148LL * (per_sm > 8 ? 8 : (per_sm < 1 ? 1 : per_sm));
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, LineTile<N>::kThreads, smem, stream>>>(n_lines, ld, mid, st);
  return cudaGetLastError();
}

// Dispatch a runtime length onto the compile-time kernels.
#define PDEOPT_LF_DISPATCH(n, CALL)                                  \
  switch (n) {                                                       \
    case 8: { constexpr int LFN = 8; CALL; } break;                  \
    case 16: { constexpr int LFN = 16; CALL; } break;                \
    case 32: { constexpr int LFN = 32; CALL; } break;                \
    case 64: { constexpr int LFN = 64; CALL; } break;                \
    case 128: { constexpr int LFN = 128; CALL; } break;              \
    case 256: { constexpr int LFN = 256; CALL; } break;              \
    case 512: { constexpr int LFN = 512; CALL; } break;              \
    default: return cudaErrorInvalidValue;                           \
  }

}  // namespace pdeopt
