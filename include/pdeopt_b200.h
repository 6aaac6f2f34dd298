/* pdeopt_b200 — C ABI of the B200-native time-stepping hot path of acoh64/pde-opt.
 *
 * The reference has no FFI: its "plugin API" is two Python protocols (a diffrax
 * AbstractSolver and an equation dataclass).  Each entry point below names the reference
 * code it replaces (paths relative to the reference root):
 *
 *   pdeopt_sifs_step_batched      K fused calls of SemiImplicitFourierSpectral.step
 *                                 (pde_opt/numerics/solvers.py:56-70) with
 *                                 CahnHilliard2DPeriodic.rhs_fd (equations/cahn_hilliard.py:89-109)
 *                                 or AllenCahn2DPeriodic.rhs_fd (equations/allen_cahn.py:81-84)
 *                                 as the vector field, i.e. the body of the diffeqsolve loop
 *                                 driven from pde_env.py:293-303 / pde_model.py:120-134,
 *                                 plus the observation / reward callbacks of
 *                                 PDEEnv.step (pde_env.py:305-309) as a fused epilogue.
 *   pdeopt_sifs_step_batched_host same, host buffers in / out (H2D + D2H inside the call).
 *   pdeopt_ad_rollout_fwd / _bwd  K fused steps of the recovered advection-diffusion equation and the
 *                                 hand-written adjoint of the rollout (custom_vjp; replaces diffrax's
 *                                 reverse-mode through diffeqsolve, pde_model.py:226-323).
 *   pdeopt_strang_step_batched    K fused calls of StrangSplitting.step (solvers.py:99-122) with
 *                                 GPE2DTSControl.B_terms (equations/gross_pitaevskii.py:67-75).
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Pointers named *_dev are CUDA
 * device pointers on the current device, *_host are host pointers.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Every function returns a
 * pdeopt_status and never throws; pdeopt_last_error() describes the last failure on the
 * calling thread.  The device-pointer entry points of fd plans never allocate; the host-buffer entry
 * point and derivs='fourier' plans grow plan-owned scratch on first use (and when the batch grows)
 * and reuse it afterwards.  A plan may be used from one thread at a time.  All device pointers of
 * one call, the stream and the plan's scratch belong to the CURRENT device: callers that use
 * several GPUs make the right device current around each call (the Python wrappers do).  There is no CPU fallback: without a CUDA device every compute entry
 * point returns PDEOPT_ERR_CUDA.
 */
#ifndef PDEOPT_B200_H
#define PDEOPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDEOPT_ABI_VERSION 4
#define PDEOPT_MAX_COEF 16
#define PDEOPT_MAX_FUSED_STEPS 512 /* per launch; callers loop for longer rollouts */
#define PDEOPT_NCTRL 8            /* floats per environment in the control block */

typedef enum {
  PDEOPT_OK = 0,
  PDEOPT_ERR_INVALID = 1,     /* bad argument */
  PDEOPT_ERR_UNSUPPORTED = 2, /* valid request the library has no kernel for */
  PDEOPT_ERR_CUDA = 3         /* CUDA runtime failure (incl. no device) */
} pdeopt_status;

typedef enum {
  PDEOPT_CH2D = 0, /* CahnHilliard2DPeriodic, cahn_hilliard.py:30-109 */
  PDEOPT_AC2D = 1, /* AllenCahn2DPeriodic,    allen_cahn.py:26-84    */
  PDEOPT_AD2D = 2, /* advection-diffusion (recovered)      */
  PDEOPT_GPE2D = 3 /* GPE2DTSControl,         gross_pitaevskii.py:18-81 */
} pdeopt_kind;

typedef enum { PDEOPT_DERIVS_FD = 0, PDEOPT_DERIVS_FOURIER = 1 } pdeopt_derivs;

/* Pointwise closure families for mu_h(c). */
typedef enum {
  PDEOPT_MU_DOUBLE_WELL = 0,     /* c^3 - c                          (tests/test_solvers.py:36) */
  PDEOPT_MU_LOG = 1,             /* log(c/(1-c)) + w(1-2c), w=coef[0] (notebooks/optimize_nn_script.py:33) */
  PDEOPT_MU_LEGENDRE = 2,        /* P(2c-1)                          (functions/legendre.py:56-74, no prior) */
  PDEOPT_MU_LEGENDRE_LOGPRIOR = 3 /* P(2c-1) + log(c/(1-c))          (same, prior_fn = log(x/(1-x))) */
} pdeopt_mu_family;

/* Pointwise closure families for the mobility D(c) / reaction rate R(c). */
typedef enum {
  PDEOPT_MOB_CONST = 0,       /* coef[0]            (ones_like, 0.15*ones) */
  PDEOPT_MOB_DEGENERATE = 1,  /* c(1-c)                                   */
  PDEOPT_MOB_ONE_PLUS_SQ = 2, /* 1 + c^2            (tests/test_rhs_convergence.py:22,55) */
  PDEOPT_MOB_LEGENDRE_EXP = 3 /* exp(P(2c-1))       (functions/legendre.py:37-53) */
} pdeopt_mob_family;

/* Control block layout, PDEOPT_NCTRL floats per environment (our definition of the
 * user callbacks update_control_value / update_control_parameter, pde_env.py:274-286):
 *   [0] mu scalar offset added to coef[0] of the mu family (e.g. the interaction w)
 *   [1] amplitude, [2] x0, [3] y0, [4] width s of an additive Gaussian bump in mu:
 *       amp * exp(-((x-x0)^2 + (y-y0)^2) / (2 s^2)), x/y the cell-centred axes
 *   [5..7] reserved (0). */

typedef struct {
  int32_t kind;   /* pdeopt_kind */
  int32_t derivs; /* pdeopt_derivs */
  int32_t nx, ny; /* collocation points per axis (axis 0 = x = slow index) */
  double lo_x, lo_y; /* lower box bounds (for the cell-centred axes used by the control bump) */
  double hx, hy;  /* grid spacings, Domain.dx (domains.py:29-32) */
  double kappa;   /* gradient-energy coefficient */
  int32_t mu_family, mu_ncoef;
  double mu_coef[PDEOPT_MAX_COEF];
  int32_t mob_family, mob_ncoef;
  double mob_coef[PDEOPT_MAX_COEF];
} pdeopt_plan_desc;

typedef struct pdeopt_plan pdeopt_plan;

int pdeopt_abi_version(void);
const char* pdeopt_last_error(void);

/* Validates the description and selects kernels.  Needs no device. */
pdeopt_status pdeopt_plan_create(const pdeopt_plan_desc* desc, pdeopt_plan** out);
pdeopt_status pdeopt_plan_destroy(pdeopt_plan* plan);

/* Number of floats of the folded symbol table for this plan: (nx/2+1)*(ny/2+1).
 * Entry [i][j] = A * fourier_symbol[i][j] for i <= nx/2, j <= ny/2 (A = the solver's splitting
 * constant, solvers.py:34,62).  The symbol of the SIFS-compatible equations is real and even in
 * each wavenumber, so one quadrant determines it.  The kernel forms 1/(1 + dt*A*symbol) per step
 * from the step's own dt, because diffrax's float32 time accumulation makes dt differ by ulps
 * from step to step. */
int64_t pdeopt_table_len(const pdeopt_plan* plan);

/* K = ksteps fused IMEX steps on `batch` independent environments.
 *   y0_dev, y1_dev : [batch][nx][ny] float32 (may alias)
 *   dt_host        : [ksteps] step lengths (t1 - t0 of each step, solvers.py:58)
 *   symbol_dev     : [pdeopt_table_len] folded A*symbol table
 *   ctrl_dev       : [batch][PDEOPT_NCTRL] control block or NULL
 *   obs_dev        : [batch][nx][ny] uint8 observation of y1 or NULL:
 *                    rint(clamp((y1-obs_lo)/(obs_hi-obs_lo),0,1)*255)   (pde_env.py:118-126)
 *   reward_dev     : [batch][2] (mean, variance) of y1 or NULL          (pde_env.py:309)
 */
pdeopt_status pdeopt_sifs_step_batched(pdeopt_plan* plan, const float* y0_dev, float* y1_dev, int32_t batch,
                                       int32_t ksteps, const float* dt_host, const float* symbol_dev,
                                       const float* ctrl_dev, uint8_t* obs_dev, float obs_lo, float obs_hi,
                                       float* reward_dev, void* stream);

/* Per-environment failure flags.  The reference has diffrax's `throw` switch only (pde_model.py:131
 * returns NaN states silently with throw=False; PDEEnv.step raises, pde_env.py:293-303); a batched
 * stepper needs to know WHICH environment failed.  After every stepping call on `plan`
 * (pdeopt_sifs_step_batched, pdeopt_sifs_filter_batched) flags_dev[b] = 1 if y1 of environment b
 * holds a NaN or Inf, else 0 (written by the fused 128x128 kernel's epilogue; one extra streaming
 * launch for the other kernels).  flags_dev: caller-owned [batch] int32 or NULL to detach. */
pdeopt_status pdeopt_plan_set_nonfinite_flags(pdeopt_plan* plan, int32_t* flags_dev);

/* The same test on any batched float32 state (Strang / GPE [batch][nx][ny][2], 3-D fields):
 * flags_dev[b] = 1 if y_dev[b*n_per_env .. (b+1)*n_per_env) holds a NaN or Inf. */
pdeopt_status pdeopt_nonfinite_flags(const float* y_dev, int32_t batch, int64_t n_per_env, int32_t* flags_dev,
                                     void* stream);

/* eq.rhs(state, t) for `batch` states (CahnHilliard2DPeriodic.rhs_fd, cahn_hilliard.py:89-109;
 * AllenCahn2DPeriodic.rhs_fd, allen_cahn.py:81-84): f_dev[b] = rhs(y_dev[b]).  May alias. */
pdeopt_status pdeopt_rhs_batched(pdeopt_plan* plan, const float* y_dev, float* f_dev, int32_t batch,
                                 const float* ctrl_dev, void* stream);

/* eq.rhs(state, t) of the finite-difference Cahn-Hilliard / Allen-Cahn equation of `plan` with the homogeneous
 * chemical potential mu_h(u) — and optionally the mobility D(u) / rate R(u) — evaluated by the CALLER on the whole
 * batch: the unfused-but-batched path for closures that are not pointwise families, i.e. the reference's neural
 * closures PeriodicCNN / Mixer2d (pde_opt/numerics/functions/cnn.py:46-102, mixer_mlp.py:40-86) or any other
 * callable.  The stencils of cahn_hilliard.py:89-109 / allen_cahn.py:81-84 run here; feed f_dev to
 * pdeopt_sifs_filter_batched for the step.  The plan's mu family is ignored; its mobility family is used when
 * mob_dev is NULL.
 *   u_dev, muh_dev, mob_dev (or NULL), f_dev : [batch][nx][ny];  work_dev : 2 * batch * nx * ny floats */
pdeopt_status pdeopt_rhs_given_mu_batched(pdeopt_plan* plan, const float* u_dev, const float* muh_dev,
                                          const float* mob_dev, float* f_dev, int32_t batch, float* work_dev, void* stream);

/* Discrete adjoint of one semi-implicit step whose mu_h (and optionally mobility) was evaluated by the caller
 * (pdeopt_rhs_given_mu_batched + pdeopt_sifs_filter_batched), up to the closure's own vector-Jacobian product: with
 * w = dt G lam1 the kernels return
 *   mubar = d loss / d mu_h  (CH: div(D_f grad_f w); AC: -D w),   dbar = d loss / d D  (CH: -1/2 sum grad_f w grad_f mu; AC: -w mu),
 *   lam0_base = lam1 - kappa lap(mubar),
 * and the caller adds  J_{mu_h}(u)^T mubar + J_D(u)^T dbar  (back-propagation through the network in its own framework;
 * the same two cotangents give the gradients of the network's parameters).  This is what jax.grad through
 * diffeqsolve does for the reference's neural closures (docs/notebooks/optimization_neural_network.ipynb).
 *   u_dev, muh_dev, mob_dev (or NULL: the plan's mobility family), lam1_dev, lam0_base_dev, mubar_dev, dbar_dev : [batch][nx][ny]
 *   work_dev : 4 * batch * nx * ny floats */
pdeopt_status pdeopt_phasefield_adjoint_given_mu(pdeopt_plan* plan, const float* u_dev, const float* muh_dev, const float* mob_dev,
                                                 const float* lam1_dev, float* lam0_base_dev, float* mubar_dev, float* dbar_dev,
                                                 int32_t batch, float dt, const float* symbol_dev, float* work_dev, void* stream);

/* One SemiImplicitFourierSpectral.step (solvers.py:56-70) with an externally evaluated vector
 * field f0 = terms.vf(t0, y0, args) (the unfused path for mu/D closures outside the enumerated
 * families): y1 = y0 + dt * Re ifft( fft(f0) / (1 + dt*A*symbol) ). */
pdeopt_status pdeopt_sifs_filter_batched(pdeopt_plan* plan, const float* y0_dev, const float* f0_dev, float* y1_dev,
                                         int32_t batch, float dt, const float* symbol_dev, void* stream);

/* Discrete adjoint of ONE semi-implicit step of the finite-difference Cahn-Hilliard / Allen-Cahn
 * equation of `plan` (no control forcing): what reverse-mode differentiation through diffeqsolve
 * yields for PDEModel.mse (pde_model.py:274-323) when the optimised leaves are the closure
 * coefficients (docs/notebooks/optimization_3D.ipynb fits Legendre coefficients of mu and D).
 *   u_dev     : [batch][nx][ny] state at the START of the step
 *   lam1_dev  : cotangent of the state after the step;  lam0_dev: cotangent before it (may alias)
 *   work_dev  : pdeopt_phasefield_adjoint_work_floats(plan, batch) floats of scratch
 *   gmu_dev, gmob_dev : [batch][PDEOPT_MAX_COEF] float64 cotangents of mu_coef / mob_coef, ACCUMULATED (+=)
 * The plan's coefficients are the point of linearisation: re-create the plan when they change. */
int64_t pdeopt_phasefield_adjoint_work_floats(const pdeopt_plan* plan, int32_t batch);
pdeopt_status pdeopt_phasefield_adjoint_step(pdeopt_plan* plan, const float* u_dev, const float* lam1_dev, float* lam0_dev,
                                             int32_t batch, float dt, const float* symbol_dev, float* work_dev,
                                             double* gmu_dev, double* gmob_dev, void* stream);

/* K fused steps like pdeopt_sifs_step_batched (no control, no epilogue) that also keep the state at the START of
 * every save_every-th step: traj_dev [ceil(ksteps / save_every)][batch][nx][ny].  This is the forward half of the
 * differentiable rollout (what diffrax's adjoints keep or recompute, pde_model.py:120-134 with
 * RecursiveCheckpointAdjoint / ForwardMode): save_every = 1 feeds pdeopt_phasefield_adjoint_step /
 * pdeopt_phasefield_tangent_steps directly; save_every = C > 1 keeps checkpoints from which the caller re-runs
 * segments of C steps with save_every = 1 during the backward sweep.  Any ksteps (the library loops over launches
 * of at most PDEOPT_MAX_FUSED_STEPS); the fused 128 x 128 kernel writes the checkpoints from shared memory. */
pdeopt_status pdeopt_sifs_rollout_fwd(pdeopt_plan* plan, const float* y0_dev, float* y1_dev, int32_t batch,
                                      int32_t ksteps, const float* dt_host, const float* symbol_dev, float* traj_dev,
                                      int32_t save_every, void* stream);

/* Discrete adjoint of the ksteps steps whose start states pdeopt_sifs_rollout_fwd kept (save_every = 1): the
 * backward half of the differentiable rollout (the custom_vjp; replaces reverse-mode differentiation through
 * diffeqsolve, pde_model.py:226-323).  128 x 128 grids run ONE fused kernel per 512 steps — the cotangent stays on
 * chip, u of every step is the only HBM read, the coefficient cotangents are reduced in the CTA; other grids loop over
 * pdeopt_phasefield_adjoint_step.
 *   traj_dev : [ksteps][batch][nx][ny];  lam1_dev: cotangent after the last step;  lam0_dev: before the first (may alias)
 *   work_dev : pdeopt_phasefield_adjoint_work_floats floats (may be NULL for 128 x 128 grids)
 *   gmu_dev, gmob_dev : [batch][PDEOPT_MAX_COEF] float64, ACCUMULATED (+=) */
pdeopt_status pdeopt_sifs_rollout_bwd(pdeopt_plan* plan, const float* traj_dev, const float* lam1_dev, float* lam0_dev,
                                      int32_t batch, int32_t ksteps, const float* dt_host, const float* symbol_dev,
                                      float* work_dev, double* gmu_dev, double* gmob_dev, void* stream);

/* Forward-mode tangent (JVP) of ksteps semi-implicit steps of the finite-difference Cahn-Hilliard / Allen-Cahn
 * equation of `plan` for ndir directions at once: the derivative optimistix's Levenberg-Marquardt takes through
 * diffrax's ForwardMode adjoint in PDEModel.train(method="least_squares") (pde_model.py:334,404-428).
 *   traj_dev : [ksteps][batch][nx][ny] state at the start of every step (pdeopt_sifs_rollout_fwd, save_every = 1)
 *   v_dev    : [ndir][batch][nx][ny] tangent of the state, advanced IN PLACE by the ksteps steps
 *   dmu_dev, dmob_dev : [ndir][PDEOPT_MAX_COEF] float32 directions in mu_coef / mob_coef (zeros for state-only tangents)
 *   work_dev : pdeopt_phasefield_tangent_work_floats(plan, batch, ndir) floats of scratch
 * The plan's coefficients are the point of linearisation. */
int64_t pdeopt_phasefield_tangent_work_floats(const pdeopt_plan* plan, int32_t batch, int32_t ndir);
pdeopt_status pdeopt_phasefield_tangent_steps(pdeopt_plan* plan, const float* traj_dev, float* v_dev, int32_t batch,
                                              int32_t ndir, int32_t ksteps, const float* dt_host, const float* dmu_dev,
                                              const float* dmob_dev, const float* symbol_dev, float* work_dev, void* stream);

/* ---- smoothed-boundary equations (cahn_hilliard.py:203-289, allen_cahn.py:87-159) ------------------------------- */
typedef struct {
  int32_t kind;   /* PDEOPT_CH2D or PDEOPT_AC2D */
  int32_t nx, ny;
  double hx, hy, kappa;
} pdeopt_sbm_desc;

/* eq.rhs(state, t) of CahnHilliard2DSmoothedBoundary.rhs_fd (cahn_hilliard.py:261-289) /
 * AllenCahn2DSmoothedBoundary.rhs_fd (allen_cahn.py:142-159) for `batch` states.  The reference integrates these
 * with explicit diffrax solvers, so there is no fused step: an RHS entry point.  The closures f, mu and D / R are
 * arbitrary callables in the reference; the caller evaluates them on the whole batch.
 *   u_dev, f_dev (free-energy density f(u)), mu_dev (mu(u)), mob_dev (D(u) or R(u)), out_dev : [batch][nx][ny]
 *   psi_dev (domain.geometry.smooth), ngp_dev (|grad psi| / psi, centred differences), side_dev (the `left_half`
 *   mask of the contact-angle term) : [nx][ny]
 *   cos_theta = cos(theta(t)), cos_pi_minus_theta = cos(pi - theta(t)) (Cahn-Hilliard only), flux = flux(t) (CH only)
 *   work_dev : batch * nx * ny floats (Cahn-Hilliard; may be NULL for Allen-Cahn) */
pdeopt_status pdeopt_sbm_rhs_batched(const pdeopt_sbm_desc* desc, const float* u_dev, const float* f_dev, const float* mu_dev,
                                     const float* mob_dev, const float* psi_dev, const float* ngp_dev, const float* side_dev,
                                     float cos_theta, float cos_pi_minus_theta, float flux, float* work_dev, float* out_dev,
                                     int32_t batch, void* stream);

/* GPE2DTSControl (gross_pitaevskii.py:18-81) geometry and constants. */
typedef struct {
  int32_t nx, ny;
  double lo_x, lo_y, hx, hy; /* Domain box lower bounds and spacings; dx of solvers.py:111 is hx */
  double k;                  /* interaction strength */
  double e;                  /* trap ellipticity */
  double trap_factor;
} pdeopt_gpe_desc;

/* K = ksteps fused calls of StrangSplitting.step (solvers.py:99-122) with GPE2DTSControl.B_terms
 * (gross_pitaevskii.py:67-75) evaluated at y0 inside the kernel, on `batch` wavefunctions.
 *   y0_dev, y1_dev : [batch][nx][ny][2] float32 (re, im) (may alias)
 *   a_term_dev     : [(nx/2+1)*(ny/2+1)][2] float32, the (kx>=0, ky>=0) quadrant of the complex
 *                    A_term (must be even in each wavenumber), or NULL when A_term is identically
 *                    zero as shipped (gross_pitaevskii.py:62) — the FFT round trips are then skipped
 *   ts_re, ts_im   : time_scale (solvers.py:90), e.g. (0,-1) for imaginary time
 *   ctrl_dev       : [batch][PDEOPT_NCTRL] or NULL; [1] amp [2] x0 [3] y0 [4] width of a Gaussian
 *                    `lights` spot amp*exp(-((x-x0)^2+(y-y0)^2)/(2 width^2)) (gross_pitaevskii.py:61,72) */
pdeopt_status pdeopt_strang_step_batched(const pdeopt_gpe_desc* desc, const float* y0_dev, float* y1_dev,
                                         int32_t batch, int32_t ksteps, const float* dt_host,
                                         const float* a_term_dev, float ts_re, float ts_im, const float* ctrl_dev,
                                         void* stream);

/* Quantised-vortex detection on batched GPE states: integer phase circulation of every grid cell
 * (pde_opt/rl_utils.py:19-84, detect_vortices; the reward helper of the GPE environments).
 *   psi_dev     : [batch][n0][n1][2] float32 (re, im)
 *   amp_thresh  : cells whose corner-averaged density is below it are suppressed (0 = off)
 *   tol         : keep windings with |circulation| >= tol * 2 pi
 *   winding_dev : [batch][n0][n1] int32 or NULL
 *   counts_dev  : [batch][3] int32 = (num_vortices, total_topological_charge, abs_charge_count) */
pdeopt_status pdeopt_gpe_detect_vortices(const float* psi_dev, int32_t batch, int32_t n0, int32_t n1,
                                         float amp_thresh, float tol, int32_t* winding_dev, int32_t* counts_dev,
                                         void* stream);

/* Advection-diffusion (recovered equation; notebooks/run_advection_diffusion.ipynb
 * cells 0-2): du/dt = -div(v u) + D lap(u), spectral derivatives, v = p0 grad exp(-r^2/(2 p1))
 * about a controlled centre, stepped by SemiImplicitFourierSpectral.step (solvers.py:56-70). */
typedef struct {
  int32_t nx, ny;
  double lo_x, lo_y, hx, hy;
} pdeopt_ad_desc;

/* Floats in the spectral table block of an advection-diffusion rollout:
 *   [ (nx/2+1)*(ny/2+1) ]  A * fourier_symbol quadrant (IMEX denominator, solvers.py:62)
 *   [ (nx/2+1)*(ny/2+1) ]  D (2 pi)^2 |k|^2 quadrant = -(D * two_pi_i_k_2).real (explicit diffusion)
 *   [ nx ] imag(two_pi_i_kx) along axis 0 with the Nyquist entry set to 0
 *   [ ny ] imag(two_pi_i_ky) along axis 1 with the Nyquist entry set to 0 */
int64_t pdeopt_ad_tables_len(const pdeopt_ad_desc* desc);

/* K = ksteps fused IMEX steps of the advection-diffusion equation (the diffeqsolve loop body
 * driven from pde_env.py:293-303 / pde_model.py:120-134) on `batch` environments.
 *   ctrl_dev : [batch][nseg][4] (cx, cy, p0, p1): velocity parameters, piecewise constant over
 *              `hold` numeric steps; local step k uses segment min((step0 + k)/hold, nseg-1)
 *   traj_dev : NULL, or where the state at the START of local step k is saved for the adjoint, at
 *              traj_dev + k*traj_stride (traj_stride in floats, >= 2*ceil(batch/2)*nx*ny).  The
 *              layout inside a step is internal (the kernels' register arrangement, per pair of
 *              environments); only pdeopt_ad_rollout_bwd reads it.                              */
pdeopt_status pdeopt_ad_rollout_fwd(const pdeopt_ad_desc* desc, const float* y0_dev, float* y1_dev, int32_t batch,
                                    int32_t ksteps, const float* dt_host, const float* tables_dev,
                                    const float* ctrl_dev, int32_t nseg, int32_t hold, int32_t step0,
                                    float* traj_dev, int64_t traj_stride, void* stream);

/* Discrete adjoint of the same K steps (the hand-written custom_vjp of the rollout; replaces
 * reverse-mode differentiation through diffeqsolve, pde_model.py:226-323):
 *   lam1_dev  : [batch][nx][ny] cotangent of the state after the K steps
 *   lam0_dev  : [batch][nx][ny] cotangent of the state before them (may alias lam1_dev)
 *   gctrl_dev : [batch][nseg][4] cotangent of ctrl_dev, ACCUMULATED (+=); zero it before the
 *               first call of a backward sweep                                                 */
pdeopt_status pdeopt_ad_rollout_bwd(const pdeopt_ad_desc* desc, const float* traj_dev, int64_t traj_stride,
                                    const float* lam1_dev, float* lam0_dev, int32_t batch, int32_t ksteps,
                                    const float* dt_host, const float* tables_dev, const float* ctrl_dev,
                                    int32_t nseg, int32_t hold, int32_t step0, float* gctrl_dev, void* stream);

/* ---- line-FFT engine (replaces jnp.fft.fftn / ifftn for fields that do not fit one SM; call sites
 * cahn_hilliard.py:156-157, gross_pitaevskii.py:58-59, solvers.py:63,:107-114) ------------------ */

/* Addressing of a batch of lines, strides in ELEMENTS of the addressed array:
 *   offset(line, idx) = (line / n_inner) * outer + (line % n_inner) * inner
 *                     + (idx / chunk) * hi + (idx % chunk) * lo
 * chunk == n gives plain strided lines; chunk < n is the packed layout of the slab all-to-all. */
typedef struct {
  int64_t n_lines, n_inner, outer, inner;
  int32_t chunk, reserved;
  int64_t hi, lo;
} pdeopt_line_geom;

/* Frequency index held at storage position `pos` after a forward transform of length n (the
 * transforms leave spectra in digit-reversed position order; -1 on bad arguments). */
int32_t pdeopt_fft_pos_to_freq(int32_t n, int32_t pos);

/* n-point transforms (n a power of two in [8, 512]) of gin->n_lines lines: forward (natural in,
 * position order out) or inverse (position order in, natural out, unnormalised), times `scale`.
 * in_dev is complex64, or float32 when in_real != 0; out_dev is complex64.  May run in place. */
pdeopt_status pdeopt_fft_lines(const void* in_dev, void* out_dev, int32_t n, const pdeopt_line_geom* gin,
                               const pdeopt_line_geom* gout, int32_t inverse, int32_t in_real, float scale, void* stream);

/* Forward transform, multiply by scale / (1 + dt * sym), inverse transform, along one axis in one
 * pass (solvers.py:62-63 on the last-transformed axis).  sym_dev: float32 A*fourier_symbol in
 * position order, addressed by gsym like the data by g. */
pdeopt_status pdeopt_fft_lines_imex(const void* in_dev, void* out_dev, int32_t n, const pdeopt_line_geom* g,
                                    const float* sym_dev, const pdeopt_line_geom* gsym, float dt, float scale,
                                    void* stream);

/* Transform fused with the slab transpose (the all-to-all of the slab-decomposed 3-D FFT): forward
 * transform of strided lines (sym_dev == NULL) or forward * scale/(1 + dt*sym) * inverse
 * (sym_dev != NULL), whose last stage stores element `pos` of every line straight into the buffer of
 * peer pos / gout->chunk (P2P stores over NVLink into peer-mapped memory, e.g. torch symmetric
 * memory), at offset src_off + line_base + (pos % chunk) * gout->lo.  peer_ptrs_host: n_peers device
 * pointers (this rank's own buffer included); gout->chunk * n_peers == n, gout->hi == 0.  The caller
 * orders the consumers with a cross-rank barrier. */
pdeopt_status pdeopt_fft_lines_to_peers(const void* in_dev, int32_t n, const pdeopt_line_geom* gin,
                                        void* const* peer_ptrs_host, int32_t n_peers, const pdeopt_line_geom* gout,
                                        int64_t src_off, const float* sym_dev, const pdeopt_line_geom* gsym, float dt,
                                        float scale, void* stream);

/* The slab transpose as a dedicated push kernel: block p (block_bytes, contiguous) of src_dev is stored into peer p's
 * buffer at dst_off_bytes (P2P stores over NVLink into peer-mapped memory; peer_ptrs_host includes this rank's own
 * buffer).  first_peer rotates the order in which the blocks are sent (pass the rank).  This is the default transport of
 * the slab-decomposed 3-D step: plain coalesced stores reach 650-700 GB/s per rank, the stores fused into the transform
 * kernels (pdeopt_fft_lines_to_peers) 270-390 GB/s.  The caller orders the consumers with a cross-rank barrier. */
pdeopt_status pdeopt_push_blocks_to_peers(const void* src_dev, void* const* peer_ptrs_host, int32_t n_peers,
                                          int64_t block_bytes, int64_t dst_off_bytes, int32_t first_peer, void* stream);

/* The same push for a PART of every block (the pipelined slab step pushes chunks while the transforms of the next
 * chunk run): for peer p, n_rows runs of run_bytes, row_stride_bytes apart, starting at src_dev + p * src_block_bytes,
 * are stored at the same row offsets from dst_off_bytes inside peer p's buffer. */
pdeopt_status pdeopt_push_rows_to_peers(const void* src_dev, void* const* peer_ptrs_host, int32_t n_peers,
                                        int64_t src_block_bytes, int64_t dst_off_bytes, int32_t n_rows, int64_t run_bytes,
                                        int64_t row_stride_bytes, int32_t first_peer, void* stream);

/* Inverse transform of the last axis fused with the update y1 = y0 + dt * Re(.) (solvers.py:63). */
pdeopt_status pdeopt_fft_lines_inv_update(const void* spec_dev, int32_t n, const pdeopt_line_geom* gin,
                                          const float* y0_dev, float* y1_dev, const pdeopt_line_geom* gout, float dt,
                                          void* stream);

/* Real contiguous lines in_dev [n_lines][n] (n_lines even) -> half spectra out_dev [n_lines][n/2+1]
 * complex64 in NATURAL frequency order h = 0..n/2 (two real lines ride in one complex transform). */
pdeopt_status pdeopt_fft_lines_r2c(const float* in_dev, void* out_dev, int32_t n, int64_t n_lines, void* stream);

/* Inverse of the above (unnormalised) fused with y1 = y0 + dt * (.) (solvers.py:63); y0/y1 [n_lines][n]. */
pdeopt_status pdeopt_fft_lines_c2r_update(const void* half_dev, int32_t n, int64_t n_lines, const float* y0_dev,
                                          float* y1_dev, float dt, void* stream);

/* ---- CahnHilliard3DPeriodic (cahn_hilliard.py:112-200) ---------------------------------------- */
typedef struct {
  int32_t nx, ny, nz; /* nx = planes held by this rank in slab mode */
  double hx, hy, hz, kappa;
  int32_t mu_family, mu_ncoef;
  double mu_coef[PDEOPT_MAX_COEF];
  int32_t mob_family, mob_ncoef;
  double mob_coef[PDEOPT_MAX_COEF];
} pdeopt_ch3d_desc;

/* rhs_fd (cahn_hilliard.py:177-200) of `batch` periodic 3-D fields u_dev [batch][nx][ny][nz] ->
 * f_dev.  Slab mode (batch == 1): halo_lo_dev [2][ny][nz] = planes x = -2, -1 and halo_hi_dev
 * [2][ny][nz] = planes nx, nx+1 of the neighbouring ranks; NULL = periodic on this rank.
 * mu_work_dev: scratch [batch][nx+2][ny][nz]. */
pdeopt_status pdeopt_ch3d_rhs(const pdeopt_ch3d_desc* desc, const float* u_dev, const float* halo_lo_dev,
                              const float* halo_hi_dev, float* mu_work_dev, float* f_dev, int32_t batch, void* stream);

int64_t pdeopt_ch3d_work_floats(const pdeopt_ch3d_desc* desc, int32_t batch);

/* ksteps calls of SemiImplicitFourierSpectral.step (solvers.py:56-70) with
 * CahnHilliard3DPeriodic.rhs_fd on `batch` whole domains (single GPU).  The field is real, so the
 * z transform is real-to-half-spectrum: symbol_pos_dev is float32 [nx][ny][nz/2+1] A*fourier_symbol
 * in position order along x and y and natural order (kz = 0..nz/2) along z;
 * work_dev: pdeopt_ch3d_work_floats. */
pdeopt_status pdeopt_ch3d_step(const pdeopt_ch3d_desc* desc, const float* y0_dev, float* y1_dev, int32_t batch,
                               int32_t ksteps, const float* dt_host, const float* symbol_pos_dev, float* work_dev,
                               void* stream);

/* Discrete adjoint of one step of pdeopt_ch3d_step (see pdeopt_phasefield_adjoint_step for the 2-D form
 * and the meaning of the arguments): docs/notebooks/optimization_3D.ipynb fits the Legendre
 * coefficients of mu and D of CahnHilliard3DPeriodic by differentiating such rollouts. */
int64_t pdeopt_ch3d_adjoint_work_floats(const pdeopt_ch3d_desc* desc, int32_t batch);
pdeopt_status pdeopt_ch3d_adjoint_step(const pdeopt_ch3d_desc* desc, const float* u_dev, const float* lam1_dev,
                                       float* lam0_dev, int32_t batch, float dt, const float* symbol_pos_dev,
                                       float* work_dev, double* gmu_dev, double* gmob_dev, void* stream);

/* ---- StrangSplitting.step on grids that do not fit one SM (256x256 complex64; any nx, ny powers of
 * two in [32, 512]): multi-kernel path on the line-FFT engine, state resident in L2.
 *   a_term_full_dev : [nx][ny][2] complex A_term in natural (fftfreq) order, or NULL when it is
 *                     identically zero (gross_pitaevskii.py:62)
 *   work_dev        : pdeopt_strang_lines_work_floats(nx, ny, batch) floats of scratch        */
int64_t pdeopt_strang_lines_work_floats(int32_t nx, int32_t ny, int32_t batch);
pdeopt_status pdeopt_strang_lines_step_batched(const pdeopt_gpe_desc* desc, const float* y0_dev, float* y1_dev,
                                               int32_t batch, int32_t ksteps, const float* dt_host,
                                               const float* a_term_full_dev, float ts_re, float ts_im,
                                               const float* ctrl_dev, float* work_dev, void* stream);

/* The same step with a caller-evaluated `lights(t, x, y)` field (gross_pitaevskii.py:61,72) added to the
 * potential: the path for light callables outside the enumerated Gaussian family (time-dependent
 * lights: one call per step with the field of that step's t0, which is where the reference evaluates
 * b, solvers.py:109).  light_dev: [nx][ny] float32 per environment, light_env_stride floats apart
 * (0 = one field shared by the batch), or NULL. */
pdeopt_status pdeopt_strang_lines_step_batched_light(const pdeopt_gpe_desc* desc, const float* y0_dev, float* y1_dev,
                                                     int32_t batch, int32_t ksteps, const float* dt_host,
                                                     const float* a_term_full_dev, float ts_re, float ts_im,
                                                     const float* ctrl_dev, const float* light_dev,
                                                     int64_t light_env_stride, float* work_dev, void* stream);

/* Same with HOST buffers: copies y0/ctrl/symbol in, runs, copies y1/obs/reward out, and
 * synchronises the stream before returning.  Scratch device memory is owned by the plan
 * (grown on first use, reused afterwards). */
pdeopt_status pdeopt_sifs_step_batched_host(pdeopt_plan* plan, const float* y0_host, float* y1_host,
                                            int32_t batch, int32_t ksteps, const float* dt_host,
                                            const float* symbol_host, const float* ctrl_host,
                                            uint8_t* obs_host, float obs_lo, float obs_hi, float* reward_host,
                                            void* stream);

/* Measured FP32 CUDA-core peak of the current device in TFLOP/s (dependent FFMA chains, 16-way
 * ILP, all SMs, best of 4 timed launches): the denominator the fused path's roofline fraction is
 * quoted against (MEASURED_PEAKS.json has no FP32 entry). */
pdeopt_status pdeopt_measure_fp32_peak(double* tflops_out, void* stream);

/* Number of kernels this library has launched in this process (bench.py "gpu_launches"). */
int64_t pdeopt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PDEOPT_B200_H */
