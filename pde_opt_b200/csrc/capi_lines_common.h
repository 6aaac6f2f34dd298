// Helpers shared by the line-FFT translation units of the C ABI.
#pragma once
#include "capi_common.h"
#include "linefft.cuh"

using namespace pdeopt;

static inline LineGeom to_geom(const pdeopt_line_geom* g) {
  LineGeom r;
  r.n_lines = g->n_lines;
  r.n_inner = g->n_inner;
  r.outer = g->outer;
  r.inner = g->inner;
  r.chunk = g->chunk;
  r.hi = g->hi;
  r.lo = g->lo;
  return r;
}
static inline bool geom_ok(const pdeopt_line_geom* g, int n) {
  return g && g->n_lines > 0 && g->n_inner > 0 && g->chunk > 0 && g->chunk <= n && (g->chunk & (g->chunk - 1)) == 0;
}
static inline bool lf_size_ok(int n) { return n >= 8 && n <= 512 && (n & (n - 1)) == 0; }

template <int MODE, bool CONTIG, class L, class M, class S>
static cudaError_t lf_run(int n, long long n_lines, L ld, M mid, S st, cudaStream_t stream) {
  PDEOPT_LF_DISPATCH(n, return (lf_launch<LFN, MODE, CONTIG>(n_lines, ld, mid, st, stream)));
  return cudaSuccess;
}

