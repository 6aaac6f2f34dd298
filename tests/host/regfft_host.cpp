// Host harness for pde_opt_b200/csrc/regfft.cuh (built by tests/test_regfft_host.py with g++).
#include "../../pde_opt_b200/csrc/regfft.cuh"
using namespace pdeopt;
template <int N, bool INV>
static void dif(float* io) {
  float2 x[N];
  for (int i = 0; i < N; ++i) x[i] = make_float2(io[2 * i], io[2 * i + 1]);
  Dif<N, 1, INV>::run(x);
  for (int i = 0; i < N; ++i) { io[2 * i] = x[i].x; io[2 * i + 1] = x[i].y; }
}
template <int N, bool INV>
static void dit(float* io) {
  float2 x[N];
  for (int i = 0; i < N; ++i) x[i] = make_float2(io[2 * i], io[2 * i + 1]);
  Dit<N, 1, INV>::run(x);
  for (int i = 0; i < N; ++i) { io[2 * i] = x[i].x; io[2 * i + 1] = x[i].y; }
}
template <int N, bool INV>
static void ditf(float* io) {
  float2 x[N];
  for (int i = 0; i < N; ++i) x[i] = make_float2(io[2 * i], io[2 * i + 1]);
  DitF<N, 1, INV>::run(x);
  for (int i = 0; i < N; ++i) { io[2 * i] = x[i].x; io[2 * i + 1] = x[i].y; }
}
extern "C" {
void ditf_fwd(int n, float* io) {
  switch (n) { case 4: ditf<4,false>(io); break; case 8: ditf<8,false>(io); break; case 16: ditf<16,false>(io); break; case 32: ditf<32,false>(io); break; }
}
void ditf_inv(int n, float* io) {
  switch (n) { case 4: ditf<4,true>(io); break; case 8: ditf<8,true>(io); break; case 16: ditf<16,true>(io); break; case 32: ditf<32,true>(io); break; }
}
void dif_fwd(int n, float* io) {
  switch (n) { case 2: dif<2,false>(io); break; case 4: dif<4,false>(io); break; case 8: dif<8,false>(io); break;
               case 16: dif<16,false>(io); break; case 32: dif<32,false>(io); break; }
}
void dit_inv(int n, float* io) {
  switch (n) { case 2: dit<2,true>(io); break; case 4: dit<4,true>(io); break; case 8: dit<8,true>(io); break;
               case 16: dit<16,true>(io); break; case 32: dit<32,true>(io); break; }
}
void dif_inv(int n, float* io) {
  switch (n) { case 4: dif<4,true>(io); break; case 8: dif<8,true>(io); break; case 16: dif<16,true>(io); break; case 32: dif<32,true>(io); break; }
}
void dit_fwd(int n, float* io) {
  switch (n) { case 4: dit<4,false>(io); break; case 8: dit<8,false>(io); break; case 16: dit<16,false>(io); break; case 32: dit<32,false>(io); break; }
}
}
