// pdeopt_jax_ffi.cc — XLA FFI (jax.ffi) custom-call handlers over the C ABI of libpdeopt_b200.so.
//
// This is the shim a pde-opt maintainer builds where jaxlib is installed (it ships the XLA FFI headers:
// `python -c "import jax.ffi; print(jax.ffi.include_dir())"`):
//
//   g++ -std=c++17 -O2 -fPIC -shared bindings/pdeopt_jax_ffi.cc -o libpdeopt_jax_ffi.so \
//       -I"$(python -c 'import jax.ffi; print(jax.ffi.include_dir())')" -Iinclude \
//       -I/usr/local/cuda/include -Lpde_opt_b200 -lpdeopt_b200 -Wl,-rpath,'$ORIGIN/../pde_opt_b200'
//
// It cannot be compiled in the image this repository was developed in (no jax / jaxlib, no network), so it is
// shipped as source; bindings/pdeopt_jax.py registers the targets and wraps them in jax.custom_vjp, and both
// are guarded on `import jax`.  Every handler is a thin argument adapter: device pointers and the XLA stream
// go straight into the entry points declared in include/pdeopt_b200.h, no copies, no allocation.
//
// Reference call sites replaced (paths relative to the reference root):
//   pdeopt_sifs_step   : the diffeqsolve loop body, SemiImplicitFourierSpectral.step (numerics/solvers.py:56-70)
//                        with CahnHilliard2DPeriodic.rhs_fd / AllenCahn2DPeriodic.rhs_fd as the vector field
//   pdeopt_strang_step : StrangSplitting.step (numerics/solvers.py:99-122) with GPE2DTSControl.B_terms
//   pdeopt_ad_fwd/_bwd : the advection-diffusion rollout and its hand-written adjoint (custom_vjp in place of
//                        reverse-mode through diffeqsolve, pde_model.py:226-323)
//   pdeopt_pf_adjoint  : discrete adjoint of one finite-difference Cahn-Hilliard / Allen-Cahn step
//   pdeopt_rollout_fwd : the forward rollout that keeps the step-start states (residuals of custom_vjp / custom_jvp)
//   pdeopt_pf_tangent  : forward-mode tangents of such steps (the ForwardMode adjoint that
//                        PDEModel.train(method="least_squares") asks diffrax for, pde_model.py:404-428)
#if __has_include("xla/ffi/api/ffi.h")
#include <cstdint>
#include <vector>

#include "xla/ffi/api/ffi.h"

#include "pdeopt_b200.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using F64 = ffi::Buffer<ffi::F64>;
using U8 = ffi::Buffer<ffi::U8>;
using Stream = ffi::PlatformStream<void*>;  // cudaStream_t, passed through as the ABI's void*

namespace {

ffi::Error Status(pdeopt_status s) {
  if (s == PDEOPT_OK) return ffi::Error::Success();
  const ffi::ErrorCode code = s == PDEOPT_ERR_INVALID ? ffi::ErrorCode::kInvalidArgument
                              : s == PDEOPT_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                            : ffi::ErrorCode::kInternal;
  return ffi::Error(code, pdeopt_last_error());
}

pdeopt_plan* Plan(int64_t handle) { return reinterpret_cast<pdeopt_plan*>(static_cast<intptr_t>(handle)); }

std::vector<float> Dts(ffi::Span<const float> dts) { return std::vector<float>(dts.begin(), dts.end()); }

// y1 = K fused IMEX steps of y0.  Operands: y0 [B, nx, ny], symbol [(nx/2+1)*(ny/2+1)] (A * fourier_symbol
// quadrant), ctrl [B, 8] (pass zeros for "no control").  Attributes: plan (address returned by
// pdeopt_plan_create, as int64), dts (the float32 step lengths of the constant-step time grid).
// Results: y1 [B, nx, ny], obs uint8 [B, nx, ny], reward [B, 2].
ffi::Error SifsStepImpl(void* stream, F32 y0, F32 symbol, F32 ctrl, int64_t plan, ffi::Span<const float> dts,
                        float obs_lo, float obs_hi, ffi::Result<F32> y1, ffi::Result<U8> obs, ffi::Result<F32> reward) {
  const std::vector<float> dt = Dts(dts);
  const int32_t batch = static_cast<int32_t>(y0.dimensions()[0]);
  return Status(pdeopt_sifs_step_batched(Plan(plan), y0.typed_data(), y1->typed_data(), batch,
                                         static_cast<int32_t>(dt.size()), dt.data(), symbol.typed_data(),
                                         ctrl.typed_data(), obs->typed_data(), obs_lo, obs_hi, reward->typed_data(),
                                         stream));
}

// One step with a caller-evaluated vector field f0 (closures outside the enumerated families, e.g. CNN / Mixer mu):
// y1 = y0 + dt * Re ifft(fft(f0) / (1 + dt * A * symbol)).
ffi::Error SifsFilterImpl(void* stream, F32 y0, F32 f0, F32 symbol, int64_t plan, float dt, ffi::Result<F32> y1) {
  return Status(pdeopt_sifs_filter_batched(Plan(plan), y0.typed_data(), f0.typed_data(), y1->typed_data(),
                                           static_cast<int32_t>(y0.dimensions()[0]), dt, symbol.typed_data(), stream));
}

// eq.rhs(state, t) of the plan's equation.
ffi::Error RhsImpl(void* stream, F32 y, F32 ctrl, int64_t plan, ffi::Result<F32> f) {
  return Status(pdeopt_rhs_batched(Plan(plan), y.typed_data(), f->typed_data(), static_cast<int32_t>(y.dimensions()[0]),
                                   ctrl.typed_data(), stream));
}

// Adjoint of one phase-field step.  Operands: u (state before the step), lam1, symbol, work (scratch of
// pdeopt_phasefield_adjoint_work_floats floats), gmu_in / gmob_in [B, 16] float64 running sums.  Results: lam0 and
// the updated sums (XLA aliases them onto the inputs through input_output_aliases on the Python side).
ffi::Error PfAdjointImpl(void* stream, F32 u, F32 lam1, F32 symbol, F32 work, F64 gmu_in, F64 gmob_in, int64_t plan,
                         float dt, ffi::Result<F32> lam0, ffi::Result<F64> gmu, ffi::Result<F64> gmob) {
  (void)gmu_in;
  (void)gmob_in;
  return Status(pdeopt_phasefield_adjoint_step(Plan(plan), u.typed_data(), lam1.typed_data(), lam0->typed_data(),
                                               static_cast<int32_t>(u.dimensions()[0]), dt, symbol.typed_data(),
                                               const_cast<float*>(work.typed_data()), gmu->typed_data(),
                                               gmob->typed_data(), stream));
}

// Forward-mode tangents of K phase-field steps for `ndir` directions at once, in place on the tangent field
// (aliased input -> output).  traj [K, B, nx, ny] are the step-start states kept by pdeopt_sifs_rollout_fwd.
ffi::Error PfTangentImpl(void* stream, F32 traj, F32 v_in, F32 dmu, F32 dmob, F32 symbol, F32 work, int64_t plan,
                         ffi::Span<const float> dts, ffi::Result<F32> v) {
  (void)v_in;  // aliased onto v by input_output_aliases
  const std::vector<float> dt = Dts(dts);
  const int32_t ndir = static_cast<int32_t>(v->dimensions()[0]);
  const int32_t batch = static_cast<int32_t>(v->dimensions()[1]);
  return Status(pdeopt_phasefield_tangent_steps(Plan(plan), traj.typed_data(), v->typed_data(), batch, ndir,
                                                static_cast<int32_t>(dt.size()), dt.data(), dmu.typed_data(), dmob.typed_data(),
                                                symbol.typed_data(), const_cast<float*>(work.typed_data()), stream));
}

// K fused steps that also keep the state at the start of every save_every-th step (the forward half of the
// differentiable rollout).  Results: y1 [B, nx, ny], traj [ceil(K / save_every), B, nx, ny].
ffi::Error RolloutFwdImpl(void* stream, F32 y0, F32 symbol, int64_t plan, ffi::Span<const float> dts, int64_t save_every,
                          ffi::Result<F32> y1, ffi::Result<F32> traj) {
  const std::vector<float> dt = Dts(dts);
  return Status(pdeopt_sifs_rollout_fwd(Plan(plan), y0.typed_data(), y1->typed_data(), static_cast<int32_t>(y0.dimensions()[0]),
                                        static_cast<int32_t>(dt.size()), dt.data(), symbol.typed_data(), traj->typed_data(),
                                        static_cast<int32_t>(save_every), stream));
}

// K fused Strang steps on 128 x 128 wavefunctions [B, n, n, 2].  a_term: the (kx >= 0, ky >= 0) quadrant of the
// complex A_term; has_a_term = 0 selects the FFT-free path of the equation as shipped (A_term identically zero).
ffi::Error StrangStepImpl(void* stream, F32 y0, F32 a_term, F32 ctrl, ffi::Span<const float> dts,
                          ffi::Span<const double> geom, float ts_re, float ts_im, int64_t has_a_term,
                          ffi::Result<F32> y1) {
  const std::vector<float> dt = Dts(dts);
  if (geom.size() != 7) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "geom = (lo_x, lo_y, hx, hy, k, e, trap_factor)");
  pdeopt_gpe_desc d{};
  d.nx = static_cast<int32_t>(y0.dimensions()[1]);
  d.ny = static_cast<int32_t>(y0.dimensions()[2]);
  d.lo_x = geom[0], d.lo_y = geom[1], d.hx = geom[2], d.hy = geom[3], d.k = geom[4], d.e = geom[5], d.trap_factor = geom[6];
  return Status(pdeopt_strang_step_batched(&d, y0.typed_data(), y1->typed_data(), static_cast<int32_t>(y0.dimensions()[0]),
                                           static_cast<int32_t>(dt.size()), dt.data(),
                                           has_a_term ? a_term.typed_data() : nullptr, ts_re, ts_im, ctrl.typed_data(),
                                           stream));
}

pdeopt_ad_desc AdDesc(const F32& y, ffi::Span<const double> geom) {
  pdeopt_ad_desc d{};
  d.nx = static_cast<int32_t>(y.dimensions()[1]);
  d.ny = static_cast<int32_t>(y.dimensions()[2]);
  d.lo_x = geom[0], d.lo_y = geom[1], d.hx = geom[2], d.hy = geom[3];
  return d;
}

// Advection-diffusion rollout; traj is the residual of the custom_vjp ([K, stride] floats, internal layout).
ffi::Error AdFwdImpl(void* stream, F32 y0, F32 tables, F32 ctrl, ffi::Span<const float> dts, ffi::Span<const double> geom,
                     int64_t hold, int64_t step0, ffi::Result<F32> y1, ffi::Result<F32> traj) {
  const std::vector<float> dt = Dts(dts);
  const pdeopt_ad_desc d = AdDesc(y0, geom);
  const int64_t stride = traj->dimensions().size() == 2 ? traj->dimensions()[1] : 0;
  return Status(pdeopt_ad_rollout_fwd(&d, y0.typed_data(), y1->typed_data(), static_cast<int32_t>(y0.dimensions()[0]),
                                      static_cast<int32_t>(dt.size()), dt.data(), tables.typed_data(), ctrl.typed_data(),
                                      static_cast<int32_t>(ctrl.dimensions()[1]), static_cast<int32_t>(hold),
                                      static_cast<int32_t>(step0), stride ? traj->typed_data() : nullptr, stride, stream));
}

ffi::Error AdBwdImpl(void* stream, F32 traj, F32 lam1, F32 tables, F32 ctrl, F32 gctrl_in, ffi::Span<const float> dts,
                     ffi::Span<const double> geom, int64_t hold, int64_t step0, ffi::Result<F32> lam0,
                     ffi::Result<F32> gctrl) {
  (void)gctrl_in;  // aliased onto gctrl by input_output_aliases: the entry point accumulates (+=)
  const std::vector<float> dt = Dts(dts);
  const pdeopt_ad_desc d = AdDesc(lam1, geom);
  return Status(pdeopt_ad_rollout_bwd(&d, traj.typed_data(), traj.dimensions()[1], lam1.typed_data(), lam0->typed_data(),
                                      static_cast<int32_t>(lam1.dimensions()[0]), static_cast<int32_t>(dt.size()),
                                      dt.data(), tables.typed_data(), ctrl.typed_data(),
                                      static_cast<int32_t>(ctrl.dimensions()[1]), static_cast<int32_t>(hold),
                                      static_cast<int32_t>(step0), gctrl->typed_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeoptSifsStep, SifsStepImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int64_t>("plan").Attr<ffi::Span<const float>>("dts")
                                  .Attr<float>("obs_lo").Attr<float>("obs_hi")
                                  .Ret<F32>().Ret<U8>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeoptSifsFilter, SifsFilterImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int64_t>("plan").Attr<float>("dt").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeoptRhs, RhsImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Attr<int64_t>("plan").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeoptPfAdjoint, PfAdjointImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F64>().Arg<F64>()
                                  .Attr<int64_t>("plan").Attr<float>("dt").Ret<F32>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeoptPfTangent, PfTangentImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int64_t>("plan").Attr<ffi::Span<const float>>("dts").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeoptRolloutFwd, RolloutFwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>()
                                  .Attr<int64_t>("plan").Attr<ffi::Span<const float>>("dts").Attr<int64_t>("save_every")
                                  .Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeoptStrangStep, StrangStepImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<ffi::Span<const float>>("dts").Attr<ffi::Span<const double>>("geom")
                                  .Attr<float>("ts_re").Attr<float>("ts_im").Attr<int64_t>("has_a_term").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeoptAdFwd, AdFwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<ffi::Span<const float>>("dts").Attr<ffi::Span<const double>>("geom")
                                  .Attr<int64_t>("hold").Attr<int64_t>("step0").Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeoptAdBwd, AdBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<ffi::Span<const float>>("dts").Attr<ffi::Span<const double>>("geom")
                                  .Attr<int64_t>("hold").Attr<int64_t>("step0").Ret<F32>().Ret<F32>());
#else
#error "pdeopt_jax_ffi.cc needs the XLA FFI headers shipped with jaxlib (jax.ffi.include_dir()); see the build line at the top"
#endif
