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

// ---- radix plan: N = R1 * R2 (* R3) -------------------------------------------------------------
// Two register stages wherever possible (radix up to 32 per thread): every stage costs a shared-memory
// round trip, a barrier and a set of address computations, and the kernels are issue / L1-bound, so
// 256 = 16 x 16 and 512 = 16 x 32 beat three radix-8 stages.
constexpr int lf_r1(int n) { return n >= 128 ? 16 : (n == 16 ? 16 : (n >= 8 ? 8 : n)); }
constexpr int lf_r2(int n) { return (n / lf_r1(n)) > 32 ? 32 : (n / lf_r1(n)); }
constexpr int lf_r3(int n) { return n / lf_r1(n) / lf_r2(n); }

// storage position -> frequency index after the forward transform
__host__ __device__ inline int line_pos_to_freq(int n, int p) {
  const int r1 = lf_r1(n), r2 = lf_r2(n);
  const int s1 = n / r1, s2 = s1 / r2;
  const int k1 = p / s1, rem = p % s1;
  const int k2 = rem / s2, k3 = rem % s2;
  return k1 + r1 * k2 + r1 * r2 * k3;
}

// Tile geometry.  CT = the kernel touches CONTIGUOUS global lines (lanes run along positions in
// its global stages); otherwise adjacent lines are adjacent in memory and lanes run across lines.
//  * An 8-byte shared-memory access is processed per half-warp, so a layout is conflict free when
//    the 16 lanes of a half-warp hit 16 distinct 8-byte slots modulo 16.
//  * T >= 16 with the odd pitch LP = T + 1: a half-warp is 16 lines at one position (16 consecutive
//    slots) or 16 consecutive positions of one line (slots 17 apart): conflict free both ways.
//  * T = 8 (512-point strided lines, halves the tile: 6 CTAs per SM instead of 3): no padding,
//    bit 0 of the position XORed with the bit that separates neighbouring last-stage blocks, so that
//    the two positions of a half-warp — neighbours, or one block apart in the last stage — always
//    land in different halves of the 16 slots.  Not usable when lanes run along positions, hence CT
//    kernels keep T = 16.
template <int N, bool CT>
struct LineTile {
  static constexpr int T = (N >= 512 && !CT) ? 8 : ((N >= 256) ? 16 : 32);  // lines per tile
  static constexpr bool kXor8 = (T == 8);
  static constexpr int LP = kXor8 ? T : T + 1;
  static constexpr int kThreads = (N >= 512 && !CT) ? 128 : 256;  // 512 = 16 x 32: the radix-32 stage has 128 work items per tile
  static constexpr size_t smem_bytes = sizeof(float2) * (size_t)(N * LP + N) + sizeof(long long) * (size_t)(3 * T);
};

// Shared-memory index of (position, line).  For the padded layouts the line index is additionally
// XOR-swizzled with the first radix digit of the position so that accesses that run along
// FREQUENCIES (consecutive h map to positions N/8 apart: the real <-> half-spectrum untangling) are
// conflict free too; accesses along lines or consecutive positions are unaffected.
template <int N, bool CT>
__device__ __forceinline__ int lf_sidx(int pos, int line) {
  if constexpr (LineTile<N, CT>::kXor8) {
    // the two positions of a half-warp are neighbours (span > radix stages) or one last-stage block
    // apart: flipping bit 0 with the bit that distinguishes neighbouring blocks separates both cases
    constexpr int LASTR = lf_r3(N) > 1 ? lf_r3(N) : (lf_r2(N) > 1 ? lf_r2(N) : lf_r1(N));
    return ((pos ^ ((pos >> ilog2(LASTR)) & 1)) << 3) + line;
  } else {
    constexpr int LP = LineTile<N, CT>::LP;
    constexpr int SH = ilog2(N / lf_r1(N));
    constexpr int MASK = (lf_r1(N) < LineTile<N, CT>::T ? lf_r1(N) : LineTile<N, CT>::T) - 1;
    return pos * LP + (line ^ ((pos >> SH) & MASK));
  }
}

// Line addressing (all strides in elements of the addressed array):
//   offset(line, idx) = (line / n_inner) * outer + (line % n_inner) * inner
//                     + (idx / chunk) * hi + (idx % chunk) * lo
// `chunk` < N expresses the packed layout of the slab all-to-all.
// Inside the kernel the two halves are tabulated in shared memory (per tile: T line bases; per
// kernel: N element offsets), so an element address costs two LDS and one add, no divisions.
struct LineGeom {
  long long n_lines;
  long long n_inner, outer, inner;
  int chunk;
  long long hi, lo;
  __host__ __device__ __forceinline__ long long line_base(long long line) const {
    return (line / n_inner) * outer + (line % n_inner) * inner;
  }
  __host__ __device__ __forceinline__ long long idx_off(int idx) const {
    return (long long)(idx / chunk) * hi + (long long)(idx % chunk) * lo;
  }
  __host__ __device__ __forceinline__ long long off(long long line, int idx) const { return line_base(line) + idx_off(idx); }
};

struct LfOff {  // element offset of position `pos` within a line: (pos / chunk) * hi + (pos % chunk) * lo, chunk = 2^shift
  int shift, mask, hi, lo;
  __device__ __forceinline__ int operator[](int pos) const { return (pos >> shift) * hi + (pos & mask) * lo; }
};

__device__ __forceinline__ LfOff lf_off(const LineGeom& g) {
  LfOff o;
  o.shift = 31 - __clz(g.chunk);  // chunk is a power of two (validated by the C ABI)
  o.mask = g.chunk - 1;
  o.hi = (int)g.hi;
  o.lo = (int)g.lo;
  return o;
}

enum : int { LF_FWD = 0, LF_INV = 1, LF_FWD_MUL_INV = 2, LF_INV_MID_FWD = 3 };

// Functor concepts (offsets are element offsets computed from the functor's own geometries):
//   Loader: LineGeom gin() ;  float2 load(long long off, long long line, int idx)
//   Mid   : LineGeom gaux();  float2 apply(float2 v, long long off_aux, long long line, int pos), split for the
//           fused forward * multiplier * inverse stage into  Aux fetch(long long off_aux, long long line)  (the
//           functor's own global reads) and  float2 apply_aux(float2 v, Aux a, long long line, int pos)
//   Storer: LineGeom gout();  float2 pre(long long off)  (the storer's own global read for that element, if any) ;
//           void store(long long off, long long line, int idx, float2 v, float2 pre) ; void flush(long long l0)
// CONTIG: elements of a line are adjacent in memory (lo == 1): warps run along idx when they touch
// global memory; otherwise adjacent lines are adjacent in memory and warps run across the T lines.
//
// The first radix stage is fused with the global loads (8 independent loads in flight per thread)
// and the last one with the global stores, so an N = R1 R2 R3 transform makes two shared-memory
// round trips instead of four; forward * multiplier * inverse shares the innermost stage in
// registers (LF_FWD_MUL_INV: five stages, four round trips).
struct LfCtx {
  float2* Sm;
  const float2* tw;
  LfOff io_in, io_out, io_aux;
  const long long *lb_in, *lb_out, *lb_aux;
  long long l0, n_lines;
};

template <int N, bool CT, int R, int S, bool INV, bool GSRC, bool GDST, bool LANE_U, class Loader, class Storer>
__device__ __forceinline__ void lf_stage(const LfCtx& c, Loader& ld, Storer& st) {
  constexpr int T = LineTile<N, CT>::T, NT = LineTile<N, CT>::kThreads;
  constexpr int SUB = S / R, NB = N / R, LOG2R = ilog2(R);
  for (int w = threadIdx.x; w < NB * T; w += NT) {
    const int u = LANE_U ? (w % NB) : (w / T), line = LANE_U ? (w / NB) : (w % T);
    const int j = u % SUB, base = (u / SUB) * S + j;
    const bool valid = c.l0 + line < c.n_lines;
    float2 x[R];
    // the storer's own global reads (e.g. y0 of the update) are issued before the transform so that their
    // latency is hidden and they are not serialised behind the stores they feed
    float2 pre[GDST ? R : 1];
    if constexpr (GDST) {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int pos = base + m * SUB;
        pre[m] = valid ? st.pre(c.lb_out[line] + c.io_out[pos]) : make_float2(0.f, 0.f);
      }
    }
    if constexpr (!INV) {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int pos = base + m * SUB;
        if constexpr (GSRC) x[m] = valid ? ld.load(c.lb_in[line] + c.io_in[pos], c.l0 + line, pos) : make_float2(0.f, 0.f);
        else x[m] = c.Sm[lf_sidx<N, CT>(pos, line)];
      }
      Dif<R, 1, false>::run(x);
      static_for<0, R>([&](auto pc) {
        constexpr int p = decltype(pc)::value;
        constexpr int k = brev<LOG2R>(p);
        float2 v = x[p];
        if constexpr (k != 0 && SUB > 1) v = cmul(v, c.tw[(j * k * (N / S)) & (N - 1)]);
        const int pos = base + k * SUB;
        if constexpr (GDST) {
          if (valid) st.store(c.lb_out[line] + c.io_out[pos], c.l0 + line, pos, v, pre[k]);
        } else {
          c.Sm[lf_sidx<N, CT>(pos, line)] = v;
        }
      });
    } else {
      static_for<0, R>([&](auto pc) {
        constexpr int p = decltype(pc)::value;
        constexpr int k = brev<LOG2R>(p);
        const int pos = base + k * SUB;
        float2 v;
        if constexpr (GSRC) v = valid ? ld.load(c.lb_in[line] + c.io_in[pos], c.l0 + line, pos) : make_float2(0.f, 0.f);
        else v = c.Sm[lf_sidx<N, CT>(pos, line)];
        if constexpr (k != 0 && SUB > 1) v = cmulc(v, c.tw[(j * k * (N / S)) & (N - 1)]);
        x[p] = v;
      });
      Dit<R, 1, true>::run(x);
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int pos = base + m * SUB;
        if constexpr (GDST) {
          if (valid) st.store(c.lb_out[line] + c.io_out[pos], c.l0 + line, pos, x[m], pre[m]);
        } else {
          c.Sm[lf_sidx<N, CT>(pos, line)] = x[m];
        }
      }
    }
  }
}

// innermost stage of forward * multiplier * inverse (span == radix, no twiddles), in registers
template <int N, bool CT, int R, bool GSRC, bool GDST, bool LANE_U, class Loader, class Mid, class Storer>
__device__ __forceinline__ void lf_stage_fmi(const LfCtx& c, Loader& ld, Mid& mid, Storer& st) {
  constexpr int T = LineTile<N, CT>::T, NT = LineTile<N, CT>::kThreads;
  constexpr int NB = N / R, LOG2R = ilog2(R);
  for (int w = threadIdx.x; w < NB * T; w += NT) {
    const int u = LANE_U ? (w % NB) : (w / T), line = LANE_U ? (w / NB) : (w % T);
    const int base = u * R;
    const bool valid = c.l0 + line < c.n_lines;
    // the multiplier's own global reads are issued first, so that their latency overlaps the loads and
    // the forward butterflies instead of sitting between the two transforms
    typename Mid::Aux aux[R];
    static_for<0, R>([&](auto pc) {
      constexpr int p = decltype(pc)::value;
      constexpr int k = brev<LOG2R>(p);
      aux[p] = valid ? mid.fetch(c.lb_aux[line] + c.io_aux[base + k], c.l0 + line) : typename Mid::Aux{};
    });
    float2 x[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if constexpr (GSRC) x[m] = valid ? ld.load(c.lb_in[line] + c.io_in[base + m], c.l0 + line, base + m) : make_float2(0.f, 0.f);
      else x[m] = c.Sm[lf_sidx<N, CT>(base + m, line)];
    }
    Dif<R, 1, false>::run(x);
    static_for<0, R>([&](auto pc) {
      constexpr int p = decltype(pc)::value;
      constexpr int k = brev<LOG2R>(p);
      if (valid) x[p] = mid.apply_aux(x[p], aux[p], c.l0 + line, base + k);
    });
    Dit<R, 1, true>::run(x);
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if constexpr (GDST) {
        if (valid) st.store(c.lb_out[line] + c.io_out[base + m], c.l0 + line, base + m, x[m], st.pre(c.lb_out[line] + c.io_out[base + m]));
      } else {
        c.Sm[lf_sidx<N, CT>(base + m, line)] = x[m];
      }
    }
  }
}

template <int N, int MODE, bool CONTIG, class Loader, class Mid, class Storer>
__global__ void __launch_bounds__(LineTile<N, CONTIG>::kThreads) linefft_kernel(long long n_lines, Loader ld, Mid mid, Storer st) {
  extern __shared__ __align__(16) unsigned char lf_smem[];
  constexpr bool CT = CONTIG;
  constexpr int T = LineTile<N, CT>::T, LP = LineTile<N, CT>::LP, NT = LineTile<N, CT>::kThreads;
  constexpr int R1 = lf_r1(N), R2 = lf_r2(N), R3 = lf_r3(N);
  constexpr int NS = 1 + (R2 > 1 ? 1 : 0) + (R3 > 1 ? 1 : 0);
  constexpr int S2 = N / R1, S3 = N / R1 / R2;
  constexpr bool LU = CONTIG;  // lane mapping of the stages that touch global memory
  float2* Sm = reinterpret_cast<float2*>(lf_smem);
  float2* tw = Sm + N * LP;
  long long* lb_in = reinterpret_cast<long long*>(tw + N);  // [T] line bases of the current tile
  long long* lb_out = lb_in + T;
  long long* lb_aux = lb_out + T;
  const LineGeom gi = ld.gin(), go = st.gout(), ga = mid.gaux();
  const LfOff io_in = lf_off(gi), io_out = lf_off(go), io_aux = lf_off(ga);
  for (int i = threadIdx.x; i < N; i += NT) {
    float s, c;
    sincospif(-2.0f * float(i) / float(N), &s, &c);
    tw[i] = make_float2(c, s);
  }
  LfCtx c{Sm, tw, io_in, io_out, io_aux, lb_in, lb_out, lb_aux, 0, n_lines};
  for (long long tile = blockIdx.x; tile * T < n_lines; tile += gridDim.x) {
    const long long l0 = tile * T;
    c.l0 = l0;
    if (threadIdx.x < T) {
      const long long l = l0 + threadIdx.x;
      lb_in[threadIdx.x] = gi.line_base(l);
      lb_out[threadIdx.x] = go.line_base(l);
      lb_aux[threadIdx.x] = ga.line_base(l);
    }
    __syncthreads();
    auto copy_in = [&]() {  // contiguous lines -> tile, warps along idx
      for (int w0 = threadIdx.x; w0 < N * T; w0 += 4 * NT) {
        float2 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int w = w0 + q * NT, idx = w % N, line = w / N;
          v[q] = (w < N * T && l0 + line < n_lines) ? ld.load(lb_in[line] + io_in[idx], l0 + line, idx) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int w = w0 + q * NT, idx = w % N, line = w / N;
          if (w < N * T) Sm[lf_sidx<N, CT>(idx, line)] = v[q];
        }
      }
      __syncthreads();
    };
    auto copy_out = [&]() {
      __syncthreads();
      // the storer's own global reads (st.pre: e.g. psi0 of the potential step) are issued in
      // batches of 4 before the dependent stores, otherwise possible aliasing serialises them
      for (int w0 = threadIdx.x; w0 < N * T; w0 += 4 * NT) {
        float2 pre[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int w = w0 + q * NT, idx = w % N, line = w / N;
          pre[q] = (w < N * T && l0 + line < n_lines) ? st.pre(lb_out[line] + io_out[idx]) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int w = w0 + q * NT, idx = w % N, line = w / N;
          if (w < N * T && l0 + line < n_lines)
            st.store(lb_out[line] + io_out[idx], l0 + line, idx, Sm[lf_sidx<N, CT>(idx, line)], pre[q]);
        }
      }
    };
    if constexpr (MODE == LF_FWD) {
      if constexpr (!CONTIG) {
        if constexpr (NS == 1) {
          lf_stage<N, CT, R1, N, false, true, true, false>(c, ld, st);
        } else if constexpr (NS == 2) {
          lf_stage<N, CT, R1, N, false, true, false, false>(c, ld, st);
          __syncthreads();
          lf_stage<N, CT, R2, S2, false, false, true, false>(c, ld, st);
        } else {
          lf_stage<N, CT, R1, N, false, true, false, false>(c, ld, st);
          __syncthreads();
          lf_stage<N, CT, R2, S2, false, false, false, false>(c, ld, st);
          __syncthreads();
          lf_stage<N, CT, R3, S3, false, false, true, false>(c, ld, st);
        }
      } else {
        lf_stage<N, CT, R1, N, false, true, false, true>(c, ld, st);
        if constexpr (NS >= 2) {
          __syncthreads();
          lf_stage<N, CT, R2, S2, false, false, false, false>(c, ld, st);
        }
        if constexpr (NS >= 3) {
          __syncthreads();
          lf_stage<N, CT, R3, S3, false, false, false, false>(c, ld, st);
        }
        copy_out();
      }
    } else if constexpr (MODE == LF_INV) {
      if constexpr (!CONTIG) {
        if constexpr (NS == 1) {
          lf_stage<N, CT, R1, N, true, true, true, false>(c, ld, st);
        } else if constexpr (NS == 2) {
          lf_stage<N, CT, R2, S2, true, true, false, false>(c, ld, st);
          __syncthreads();
          lf_stage<N, CT, R1, N, true, false, true, false>(c, ld, st);
        } else {
          lf_stage<N, CT, R3, S3, true, true, false, false>(c, ld, st);
          __syncthreads();
          lf_stage<N, CT, R2, S2, true, false, false, false>(c, ld, st);
          __syncthreads();
          lf_stage<N, CT, R1, N, true, false, true, false>(c, ld, st);
        }
      } else {
        copy_in();
        if constexpr (NS >= 3) {
          lf_stage<N, CT, R3, S3, true, false, false, false>(c, ld, st);
          __syncthreads();
        }
        if constexpr (NS >= 2) {
          lf_stage<N, CT, R2, S2, true, false, false, false>(c, ld, st);
          __syncthreads();
        }
        lf_stage<N, CT, R1, N, true, false, true, true>(c, ld, st);
      }
    } else if constexpr (MODE == LF_INV_MID_FWD) {
      // contiguous lines: inverse transform, a pointwise operator in natural order (with its own
      // global reads, e.g. psi0 of the potential step), forward transform again, one pass over memory.
      // Mid concept here: float2 pre(long long off) ; float2 apply(float2 v, long long off, long long line,
      // int idx, float2 pre) ; void flush(long long l0)
      static_assert(MODE != LF_INV_MID_FWD || CONTIG, "LF_INV_MID_FWD is for contiguous lines");
      copy_in();
      if constexpr (NS >= 3) {
        lf_stage<N, CT, R3, S3, true, false, false, false>(c, ld, st);
        __syncthreads();
      }
      if constexpr (NS >= 2) {
        lf_stage<N, CT, R2, S2, true, false, false, false>(c, ld, st);
        __syncthreads();
      }
      lf_stage<N, CT, R1, N, true, false, false, false>(c, ld, st);
      __syncthreads();
      for (int w0 = threadIdx.x; w0 < N * T; w0 += 4 * NT) {
        float2 pre[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int w = w0 + q * NT, idx = w % N, line = w / N;
          pre[q] = (w < N * T && l0 + line < n_lines) ? mid.pre(lb_out[line] + io_out[idx]) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int w = w0 + q * NT, idx = w % N, line = w / N;
          if (w < N * T && l0 + line < n_lines) {
            const int si = lf_sidx<N, CT>(idx, line);
            Sm[si] = mid.apply(Sm[si], lb_out[line] + io_out[idx], l0 + line, idx, pre[q]);
          }
        }
      }
      __syncthreads();
      lf_stage<N, CT, R1, N, false, false, false, false>(c, ld, st);
      if constexpr (NS >= 2) {
        __syncthreads();
        lf_stage<N, CT, R2, S2, false, false, false, false>(c, ld, st);
      }
      if constexpr (NS >= 3) {
        __syncthreads();
        lf_stage<N, CT, R3, S3, false, false, false, false>(c, ld, st);
      }
      copy_out();
      mid.flush(l0);
    } else {
      if constexpr (NS == 1) {
        lf_stage_fmi<N, CT, R1, true, true, LU>(c, ld, mid, st);
      } else if constexpr (NS == 2) {
        lf_stage<N, CT, R1, N, false, true, false, LU>(c, ld, st);
        __syncthreads();
        lf_stage_fmi<N, CT, R2, false, false, false>(c, ld, mid, st);
        __syncthreads();
        lf_stage<N, CT, R1, N, true, false, true, LU>(c, ld, st);
      } else {
        lf_stage<N, CT, R1, N, false, true, false, LU>(c, ld, st);
        __syncthreads();
        lf_stage<N, CT, R2, S2, false, false, false, false>(c, ld, st);
        __syncthreads();
        lf_stage_fmi<N, CT, R3, false, false, false>(c, ld, mid, st);
        __syncthreads();
        lf_stage<N, CT, R2, S2, true, false, false, false>(c, ld, st);
        __syncthreads();
        lf_stage<N, CT, R1, N, true, false, true, LU>(c, ld, st);
      }
    }
    st.flush(l0);  // optional per-tile reduction hook (uniform; may contain barriers)
    __syncthreads();
  }
}

// ---- real <-> half-spectrum transforms of contiguous lines ---------------------------------------
// Two adjacent REAL lines a, b ride in one complex line z = a + i b (the same packing as the fused
// 2-D kernels); the forward transform is untangled into the two half spectra
//   A[h] = (Z[h] + conj Z[N-h]) / 2,   B[h] = (Z[h] - conj Z[N-h]) / (2i),   h = 0 .. N/2,
// stored in NATURAL order of h (the remaining axes stay in position order), which halves the data
// every later pass of a real field's 3-D transform has to move.  The inverse rebuilds Z from (A, B).
struct LfNoIo {
  __device__ __forceinline__ float2 load(long long, long long, int) const { return make_float2(0.f, 0.f); }
  __device__ __forceinline__ float2 pre(long long) const { return make_float2(0.f, 0.f); }
  __device__ __forceinline__ void store(long long, long long, int, float2, float2) const {}
};

template <int N>
struct LineTileReal {
  static constexpr int T = LineTile<N, true>::T;
  static constexpr size_t smem_bytes = sizeof(float2) * (size_t)(N * (T + 1) + N) + sizeof(int) * (size_t)N + sizeof(long long) * (size_t)T;
};

// Adapters that let the generic stages talk to a real-line Io: the first forward stage loads pairs of real
// lines straight from global memory, the last inverse stage stores them (offset = local pair * N + idx).
template <class Io>
struct LfRealLoad {
  Io io;
  __device__ __forceinline__ float2 load(long long, long long pair, int idx) const { return io.load_pair(pair, idx); }
};
template <int N, class Io>
struct LfRealStore {
  Io io;
  long long l0;
  __device__ __forceinline__ float2 pre(long long off) const { return io.pre_pair(l0 + (off / N), (int)(off % N)); }
  __device__ __forceinline__ void store(long long, long long pair, int idx, float2 v, float2 p) const { io.store_pair(pair, idx, v, p); }
};

// Io (forward):  float2 load_pair(long long pair, int idx) ; void store_half(long long pair, int h, float2 A, float2 B)
// Io (inverse):  void load_half(long long pair, int h, float2& A, float2& B) ; float2 pre_pair(long long pair, int idx) ; void store_pair(long long pair, int idx, float2 v, float2 pre)
template <int N, bool INV, class Io>
__global__ void __launch_bounds__(LineTile<N, true>::kThreads) linefft_real_kernel(long long n_pairs, Io io) {
  extern __shared__ __align__(16) unsigned char lf_smem[];
  constexpr bool CT = true;
  constexpr int T = LineTile<N, CT>::T, LP = LineTile<N, CT>::LP, NT = LineTile<N, CT>::kThreads;
  constexpr int R1 = lf_r1(N), R2 = lf_r2(N), R3 = lf_r3(N);
  constexpr int S2 = N / R1, S3 = N / R1 / R2, HP = N / 2 + 1;
  constexpr int LH = (N * T / NT >= 16) ? 8 : 4;  // half-spectrum loads in flight per thread
  float2* Sm = reinterpret_cast<float2*>(lf_smem);
  float2* tw = Sm + N * LP;
  int* f2p = reinterpret_cast<int*>(tw + N);  // frequency -> storage position
  long long* lb = reinterpret_cast<long long*>(f2p + N);  // [T] local pair * N: offsets handed to the storer adapter
  for (int i = threadIdx.x; i < N; i += NT) {
    float s, c;
    sincospif(-2.0f * float(i) / float(N), &s, &c);
    tw[i] = make_float2(c, s);
    f2p[line_pos_to_freq(N, i)] = i;
  }
  if (threadIdx.x < T) lb[threadIdx.x] = (long long)threadIdx.x * N;
  const LfOff ident{31, 0x7fffffff, 0, 1};  // offset of position pos = pos
  LfCtx c{Sm, tw, ident, ident, LfOff{}, lb, lb, nullptr, 0, n_pairs};
  LfNoIo nio;
  for (long long tile = blockIdx.x; tile * T < n_pairs; tile += gridDim.x) {
    const long long l0 = tile * T;
    c.l0 = l0;
    __syncthreads();
    if constexpr (!INV) {
      // first radix stage fused with the global loads of the real pairs (lanes along the positions)
      LfRealLoad<Io> rl{io};
      lf_stage<N, CT, R1, N, false, true, false, true>(c, rl, nio);
      __syncthreads();
      if constexpr (R2 > 1) {
        lf_stage<N, CT, R2, S2, false, false, false, false>(c, nio, nio);
        __syncthreads();
      }
      if constexpr (R3 > 1) {
        lf_stage<N, CT, R3, S3, false, false, false, false>(c, nio, nio);
        __syncthreads();
      }
      for (int w = threadIdx.x; w < HP * T; w += NT) {
        const int h = w % HP, pl = w / HP;
        if (l0 + pl < n_pairs) {
          const float2 zk = Sm[lf_sidx<N, CT>(f2p[h], pl)], zn = Sm[lf_sidx<N, CT>(f2p[(N - h) & (N - 1)], pl)];
          const float2 A = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
          const float2 B = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
          io.store_half(l0 + pl, h, A, B);
        }
      }
    } else {
      for (int w0 = threadIdx.x; w0 < HP * T; w0 += LH * NT) {
        float2 A[LH], B[LH];
#pragma unroll
        for (int q = 0; q < LH; ++q) {
          const int w = w0 + q * NT, h = w % HP, pl = w / HP;
          A[q] = B[q] = make_float2(0.f, 0.f);
          if (w < HP * T && l0 + pl < n_pairs) io.load_half(l0 + pl, h, A[q], B[q]);
        }
#pragma unroll
        for (int q = 0; q < LH; ++q) {
          const int w = w0 + q * NT, h = w % HP, pl = w / HP;
          if (w < HP * T) {
            // Z[h] = A + i B ;  Z[N-h] = conj(A) + i conj(B)
            Sm[lf_sidx<N, CT>(f2p[h], pl)] = make_float2(A[q].x - B[q].y, A[q].y + B[q].x);
            if (h != 0 && h != N / 2) Sm[lf_sidx<N, CT>(f2p[N - h], pl)] = make_float2(A[q].x + B[q].y, B[q].x - A[q].y);
          }
        }
      }
      __syncthreads();
      if constexpr (R3 > 1) {
        lf_stage<N, CT, R3, S3, true, false, false, false>(c, nio, nio);
        __syncthreads();
      }
      if constexpr (R2 > 1) {
        lf_stage<N, CT, R2, S2, true, false, false, false>(c, nio, nio);
        __syncthreads();
      }
      // last radix stage fused with the global stores (and the storer's own reads) of the real pairs
      LfRealStore<N, Io> rs{io, l0};
      lf_stage<N, CT, R1, N, true, false, true, true>(c, nio, rs);
    }
  }
}

template <int N, bool INV, class Io>
cudaError_t lf_launch_real(long long n_pairs, Io io, cudaStream_t stream) {
  auto kern = linefft_real_kernel<N, INV, Io>;
  constexpr size_t smem = LineTileReal<N>::smem_bytes;
  static bool attr[64] = {};  // kernel attributes are per-device state
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !attr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  const long long tiles = (n_pairs + LineTile<N, true>::T - 1) / LineTile<N, true>::T;
  const int per_sm = (int)((227 * 1024) / (smem + 1024));
  long long grid = tiles;
  const long long cap = 148LL * (per_sm > 8 ? 8 : (per_sm < 1 ? 1 : per_sm));
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, LineTile<N, true>::kThreads, smem, stream>>>(n_pairs, io);
  return cudaGetLastError();
}

// ---- stock functors ---------------------------------------------------------------------------
struct LfLoadC {  // complex array
  const float2* p;
  LineGeom g;
  __device__ __forceinline__ LineGeom gin() const { return g; }
  __device__ __forceinline__ float2 load(long long off, long long, int) const { return p[off]; }
};
struct LfStoreC {
  float2* p;
  LineGeom g;
  __device__ __forceinline__ LineGeom gout() const { return g; }
  __device__ __forceinline__ float2 pre(long long) const { return make_float2(0.f, 0.f); }
  __device__ __forceinline__ void store(long long off, long long, int, float2 v, float2) const { p[off] = v; }
  __device__ __forceinline__ void flush(long long) {}
};
struct LfStoreCScaled {
  float2* p;
  LineGeom g;
  float scale;
  __device__ __forceinline__ LineGeom gout() const { return g; }
  __device__ __forceinline__ float2 pre(long long) const { return make_float2(0.f, 0.f); }
  __device__ __forceinline__ void store(long long off, long long, int, float2 v, float2) const {
    p[off] = make_float2(v.x * scale, v.y * scale);
  }
  __device__ __forceinline__ void flush(long long) {}
};
// Fused transpose of the slab decomposition: the last stage of a transform stores straight into the
// PEER GPUs' buffers over NVLink (P2P stores into symmetric memory), the destination rank being the
// chunk index of the position, instead of writing a local send buffer for an NCCL all-to-all.
struct LfStorePeers {
  float2* peers[8];
  LineGeom g;         // geometry inside one peer's buffer (hi must be 0: the chunk selects the peer)
  int shift;          // log2(chunk)
  long long src_off;  // where this rank's block starts inside every peer's buffer
  __device__ __forceinline__ LineGeom gout() const { return g; }
  __device__ __forceinline__ float2 pre(long long) const { return make_float2(0.f, 0.f); }
  __device__ __forceinline__ void store(long long off, long long, int pos, float2 v, float2) const {
    peers[pos >> shift][src_off + off] = v;
  }
  __device__ __forceinline__ void flush(long long) {}
};
struct LfMidNone {
  struct Aux {};
  __device__ __forceinline__ LineGeom gaux() const { return LineGeom{1, 1, 0, 0, 1, 0, 0}; }
  __device__ __forceinline__ float2 apply(float2 v, long long, long long, int) const { return v; }
  __device__ __forceinline__ Aux fetch(long long, long long) const { return Aux{}; }
  __device__ __forceinline__ float2 apply_aux(float2 v, Aux, long long, int) const { return v; }
};

template <int N, int MODE, bool CONTIG, class Loader, class Mid, class Storer>
cudaError_t lf_launch(long long n_lines, Loader ld, Mid mid, Storer st, cudaStream_t stream) {
  auto kern = linefft_kernel<N, MODE, CONTIG, Loader, Mid, Storer>;
  constexpr size_t smem = LineTile<N, CONTIG>::smem_bytes;
  static bool attr[64] = {};  // kernel attributes are per-device state
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !attr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  const long long tiles = (n_lines + LineTile<N, CONTIG>::T - 1) / LineTile<N, CONTIG>::T;
  const int per_sm = (int)((227 * 1024) / (smem + 1024));
  long long grid = tiles;
  const long long cap = 148LL * (per_sm > 8 ? 8 : (per_sm < 1 ? 1 : per_sm));
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, LineTile<N, CONTIG>::kThreads, smem, stream>>>(n_lines, ld, mid, st);
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
