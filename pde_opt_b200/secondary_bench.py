"""Bounded measurements of the BASELINE configs other than the headline one (C3 GPE 256x256 Strang,
C4 advection-diffusion rollout + adjoint, C5 Cahn-Hilliard 512^3, and - under torchrun - the
slab-decomposed 512^3 step at N ranks): the `secondary` block of bench.py's JSON line.

Each entry carries its own roofline {bound, achieved, peak, unit, frac}:
  * FP32-bound kernels: algorithmic flop per env-step (5 N log2 N per complex transform, 2.5 N log2 N
    per real one, counted pointwise flops) / CUDA-event time, against the FFMA-chain peak measured live;
  * HBM-bound passes: algorithmic bytes per step / time against the measured copy bandwidth (MEASURED_PEAKS.json);
  * the slab step: the slower of (local algorithmic bytes / measured HBM) and (bytes that must cross NVLink per
    rank / the measured 770 GB/s per direction) is the target time; frac = target / measured.
Everything is timed with CUDA events after warm-up; multi-rank numbers are the max over ranks."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVLINK_GBS = 770.0  # measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md)


def _timed(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def _max_over_ranks(t):
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        x = torch.tensor([t], device="cuda", dtype=torch.float64)
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return float(x.item())
    return t


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _roof_fp32(flop, seconds, peak_tf):
    a = flop / seconds / 1e12
    return {"bound": "fp32", "achieved": a, "peak": peak_tf, "unit": "TFLOP/s", "frac": a / peak_tf if peak_tf else None}


def _roof_hbm(nbytes, seconds):
    pk, how = hbm_peak()
    a = nbytes / seconds / 1e9
    return {"bound": "hbm", "achieved": a, "peak": pk, "unit": "GB/s", "frac": a / pk, "peak_source": how}


def c3_gpe(envs=128, K=16, peak_tf=None, world=1):
    """BASELINE config 3: GPE 256x256 complex64 Strang, `envs` environments per GPU (1024 over 8 GPUs)."""
    from . import Domain
    from .equations import GPE2DTSControl
    from .solvers import ODETerm, StrangSplitting

    N = 256
    L_ = 29.4
    dom = Domain((N, N), ((-L_ / 2, L_ / 2),) * 2, "dimensionless")
    eq = GPE2DTSControl(dom, 3371.7, 0.0, None, 1.0)
    g = np.load(os.path.join(ROOT, "tests", "golden", "gpe_ground_state_256.npy"))
    rng = np.random.default_rng(int(os.environ.get("RANK", "0")))
    y = torch.from_numpy(np.stack([g * (1 + 0.01 * rng.normal(size=g.shape)) for _ in range(envs)]).astype(np.float32)).cuda()
    times = np.arange(K + 1, dtype=np.float32) * np.float32(2 * np.pi * 1e-4)
    out = []
    for kinetic in (True, False):
        a = (0.5j * eq.two_pi_i_k_2).astype(np.complex64) if kinetic else eq.A_term
        solver = StrangSplitting(a, eq.dx, eq.fft, eq.ifft, -1j)
        buf = torch.empty_like(y)
        t = _max_over_ranks(_timed(lambda: solver.rollout(ODETerm(eq), times, y, out=buf)))
        flop_pt = (4 * 5 * 16 + 40) if kinetic else 40
        out.append({
            "config": "C3 GPE 256x256 c64 Strang, imaginary time, " + ("kinetic term on (A_term = 0.5j (2 pi i k)^2)" if kinetic else "as shipped (A_term = 0)"),
            "envs_per_gpu": envs, "n_gpus": world, "fused_steps": K, "value": world * envs * K / t, "unit": "env-steps/s",
            "ms_per_env_step_batch": t / K * 1e3,
            "roofline": {**_roof_fp32(flop_pt * N * N * envs * K, t, peak_tf),
                         "how": f"{flop_pt} algorithmic flop per grid point per step (4 complex 256x256 transforms = 320, pointwise 40)"},
        })
    return out


def c4_ad(envs=512, K=500, peak_tf=None):
    """BASELINE config 4: advection-diffusion 128x128, 500-step rollout, forward and forward + hand-written adjoint."""
    from . import Domain
    from .adjoint import ad_rollout
    from .equations import AdvectionDiffusion2D
    from .functions import GaussianVelocity

    N, H = 128, 0.02
    dom = Domain((N, N), ((-N * H / 2, N * H / 2),) * 2, "dimensionless")
    eq = AdvectionDiffusion2D(dom, GaussianVelocity(0.1, 0.01), 0.1)
    rng = np.random.default_rng(0)
    y0 = torch.from_numpy((0.5 + 0.01 * rng.normal(size=(envs, N, N))).astype(np.float32)).cuda()
    ctrl = torch.tensor([0.1, -0.1, 0.1, 0.01], device="cuda").expand(envs, 10, 4).contiguous()
    times = np.arange(K + 1, dtype=np.float32) * np.float32(1e-4)
    t_f = _timed(lambda: ad_rollout(eq, y0, ctrl, times, hold=50), 1, 3)

    def fb():
        yg, cg = y0.clone().requires_grad_(True), ctrl.clone().requires_grad_(True)
        (ad_rollout(eq, yg, cg, times, hold=50) ** 2).mean().backward()

    t_fb = _timed(fb, 1, 2)
    flop = 2.5e6  # 4 real-FFT equivalents = 140 flop/pt + ~12 pointwise, per env-step forward
    torch.cuda.empty_cache()
    return [
        {"config": "C4 advection-diffusion 128x128, 500-step rollout, forward", "envs_per_gpu": envs, "steps": K,
         "value": envs * K / t_f, "unit": "env-steps/s", "roofline": {**_roof_fp32(flop * envs * K, t_f, peak_tf), "how": "2.5 MFLOP per env-step forward (SURVEY 8d)"}},
        {"config": "C4 advection-diffusion 128x128, 500-step rollout, forward + adjoint (gradients of u0 and the control path)",
         "envs_per_gpu": envs, "steps": K, "value": envs * K / t_fb, "unit": "env-steps/s", "trajectory_GiB": envs * K * N * N * 4 / 2**30,
         "roofline": {**_roof_fp32(2 * flop * envs * K, t_fb, peak_tf), "how": "forward + adjoint counted as 2 x 2.5 MFLOP per env-step"}},
    ]


def c2_variants(peak_tf=None, envs=4096):
    """The headline equation at K = 1 (the launch-per-step path PDEModel.solve with dense SaveAt, PID stepping and the
    unfused `given f0` route use; HBM view) and its differentiable rollout: fused forward that keeps every state +
    fused adjoint (pdeopt_sifs_rollout_fwd / _bwd), Legendre closures, 512 environments x 256 steps."""
    from . import Domain
    from .equations import CahnHilliard2DPeriodic
    from .functions import ChemicalPotentialLegendrePolynomials, DegenerateMobility, DiffusionLegendrePolynomials, LogRegular
    from .solvers import SemiImplicitFourierSpectral

    N, H = 128, 0.01
    dom = Domain((N, N), ((-N * H / 2, N * H / 2),) * 2, "dimensionless")
    g = torch.Generator(device="cuda").manual_seed(5)
    out = []
    # ---- K = 1 ----
    eq = CahnHilliard2DPeriodic(dom, 0.002, LogRegular(3.0), DegenerateMobility())
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    plan, sym = eq.plan(), solver.symbol_on("cuda")
    y = (0.5 + 0.01 * torch.randn((envs, N, N), device="cuda", generator=g)).clamp_(0, 1).contiguous()
    buf = torch.empty_like(y)
    dts = np.full(1, 1e-6, np.float32)
    t = _timed(lambda: plan.step(y, dts, sym, out=buf), 3, 10)
    out.append({"config": "C2 Cahn-Hilliard 128x128, K = 1 (one launch per numeric step: state read + written every step)",
                "envs_per_gpu": envs, "value": envs / t, "unit": "env-steps/s",
                "roofline": {**_roof_hbm(2 * 4 * N * N * envs, t), "how": "2*4*128*128 B per env-step (SURVEY 8d K=1 view)",
                             "fp32_frac": (108 * N * N * envs / t / 1e12) / peak_tf if peak_tf else None,
                             "note": "compute-bound even at K = 1: the step itself needs envs / (K>=16 rate) of the time"}})
    del y, buf
    # ---- forward with kept states + fused adjoint ----
    B, K = 512, 256
    eq = CahnHilliard2DPeriodic(dom, 0.002, ChemicalPotentialLegendrePolynomials([0.1, -2.2, 0.3, 0.25], "log"),
                                DiffusionLegendrePolynomials([-0.3, 0.2, -0.1]))
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    plan, sym = eq.plan(), solver.symbol_on("cuda")
    y = (0.5 + 0.05 * torch.randn((B, N, N), device="cuda", generator=g)).clamp_(0.1, 0.9).contiguous()
    dts = np.full(K, 1e-6, np.float32)
    lam = torch.randn((B, N, N), device="cuda", generator=g)
    gmu = torch.zeros((B, 16), dtype=torch.float64, device="cuda")
    gmob = torch.zeros_like(gmu)
    y1 = torch.empty_like(y)
    state = {}

    def fwd():
        state["traj"] = plan.rollout_fwd(y, dts, sym, save_every=1, out=y1)[1]

    t_f = _timed(fwd, 1, 2)
    t_b = _timed(lambda: plan.rollout_bwd(state["traj"], lam.clone(), dts, sym, gmu, gmob), 1, 2)
    f_fwd, f_bwd = 108 + 40, 70 + 95  # per grid point and step: forward with Legendre closures; adjoint = filter + 3-level stencil
    out.append({"config": "C2 Cahn-Hilliard 128x128 differentiable rollout: fused forward keeping every state + fused K-step adjoint "
                          "(Legendre mu with log prior, exp-Legendre mobility; gradients of 7 coefficients and y0)",
                "envs_per_gpu": B, "steps": K, "value": B * K / (t_f + t_b), "unit": "env-steps/s (forward + adjoint)",
                "forward_env_steps_per_s": B * K / t_f, "adjoint_env_steps_per_s": B * K / t_b,
                "trajectory_GiB": B * K * N * N * 4 / 2**30,
                "roofline": {**_roof_fp32((f_fwd + f_bwd) * N * N * B * K, t_f + t_b, peak_tf),
                             "how": f"{f_fwd} (forward) + {f_bwd} (adjoint: 2 real FFTs = 70, recomputed mu/D, flux transposes, lap, closure derivatives ~ 95) "
                                    "algorithmic flop per grid point and step",
                             "hbm_GBs": (2 * 4 * N * N * B * K) / (t_f + t_b) / 1e9}})
    del state, y, lam, y1
    torch.cuda.empty_cache()
    return out


def _ch3d_problem(n, rank=0, world=1):
    from . import Domain
    from .equations import CahnHilliard3DPeriodic
    from .functions import ConstantMobility, LogRegular
    from .linefft import pos_to_freq

    pts = (n, n, n)
    dom = Domain(pts, tuple((0.0, n * 0.01) for _ in range(3)), "dimensionless")
    eq = CahnHilliard3DPeriodic(dom, 0.002, LogRegular(3.0), ConstantMobility(0.15))
    C = n // world
    knat = (2 * np.pi * np.fft.fftfreq(n, 0.01)).astype(np.float32)
    kk = knat[pos_to_freq(n)] ** 2
    kx = torch.as_tensor(kk, device="cuda")
    ky = torch.as_tensor(kk[rank * C:(rank + 1) * C], device="cuda")
    kz = torch.as_tensor(knat[: n // 2 + 1] ** 2, device="cuda")
    k2 = (kx[:, None, None] + ky[None, :, None]) + kz[None, None, :]
    sym = (0.5 * 0.002 * k2 * k2).contiguous()  # A * kappa * |k|^4 in position order along x, y
    del k2
    return eq, sym


def c5_ch3d(n=512):
    """BASELINE config 5 on ONE GPU: Cahn-Hilliard 512^3 semi-implicit step on the line-FFT engine."""
    eq, sym = _ch3d_problem(n)
    g = torch.Generator(device="cuda").manual_seed(0)
    u = (0.5 + 0.01 * torch.randn((1, n, n, n), device="cuda", generator=g)).clamp_(0.01, 0.99)
    out = torch.empty_like(u)
    plan = eq.plan()
    K = 4
    dts = np.full(K, 1e-6, np.float32)
    t = _timed(lambda: plan.step(u, dts, sym, out=out), 1, 3) / K
    vol = n**3
    alg = (2 * 4 * vol) + 2 * 3 * 2 * (8 * (n // 2 + 1) * n * n)  # RHS r+w, 2 x three r+w passes of the half spectrum
    r = {"config": f"C5 Cahn-Hilliard 3D {n}^3, one GPU, one semi-implicit step", "value": vol / t, "unit": "grid-point-steps/s",
         "ms_per_step": t * 1e3, "roofline": {**_roof_hbm(alg, t), "how": f"{alg / 1e9:.2f} GB algorithmic per step (SURVEY 8d)"},
         "finite": bool(torch.isfinite(out).all())}
    del u, out, sym
    torch.cuda.empty_cache()
    return [r]


def c5_slab(n=512):
    """BASELINE config 5 sharded over the ranks of this job: x-slabs, two transposes per step."""
    import torch.distributed as dist

    from .parallel import SlabCahnHilliard3D

    rank, world = dist.get_rank(), dist.get_world_size()
    eq, sym = _ch3d_problem(n, rank, world)
    nxl = n // world
    slab = SlabCahnHilliard3D(eq, 0.5, device=torch.device("cuda", torch.cuda.current_device()), symbol_pos_local=sym, transport="auto")
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    u = (0.5 + 0.01 * torch.randn((nxl, n, n), device="cuda", generator=g)).clamp_(0.01, 0.99)
    out = torch.empty_like(u)
    for _ in range(2):
        slab.step(u, 1e-6, out=out)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    e0.record()
    for _ in range(iters):
        slab.step(u, 1e-6, out=out)
    e1.record()
    torch.cuda.synchronize()
    t = _max_over_ranks(e0.elapsed_time(e1) / iters * 1e-3)
    vol = n**3
    alg_local = ((2 * 4 * vol) + 2 * 3 * 2 * (8 * (n // 2 + 1) * n * n)) / world
    a2a = 2 * nxl * n * (n // 2 + 1) * 8 * (world - 1) / world  # bytes leaving each rank per step (two transposes)
    pk, _ = hbm_peak()
    target = max(alg_local / (pk * 1e9), a2a / (NVLINK_GBS * 1e9))
    finite = bool(torch.isfinite(out).all())
    transport = slab.transport
    del u, out, sym, slab
    torch.cuda.empty_cache()
    return [{"config": f"C5 Cahn-Hilliard 3D {n}^3 slab-decomposed over {world} GPUs, one semi-implicit step", "n_gpus": world,
             "value": vol / t, "unit": "grid-point-steps/s", "ms_per_step": t * 1e3, "transposed_MB_per_rank_per_step": a2a / 1e6,
             "transport": transport,
             "roofline": {"bound": "hbm+nvlink", "achieved": 1.0 / t, "peak": 1.0 / target, "unit": "steps/s", "frac": target / t,
                          "how": f"target = max(local algorithmic bytes {alg_local / 1e9:.2f} GB / measured HBM, {a2a / 1e6:.0f} MB per rank over NVLink / {NVLINK_GBS:.0f} GB/s)"},
             "finite": finite}]


def run_all(peak_tf, world=1, budget_s=120.0):
    """The `secondary` block.  On one GPU: C3, C4, C5.  Under torchrun: C3 at 128 environments per GPU and the
    slab-decomposed 512^3 step over all ranks (every rank participates; rank 0 reports)."""
    import time

    t0 = time.perf_counter()
    out = []
    steps = [lambda: c3_gpe(128, 16, peak_tf, world)]
    if world == 1:
        steps += [lambda: c2_variants(peak_tf), lambda: c4_ad(512, 500, peak_tf), lambda: c5_ch3d(512)]
    else:
        steps += [lambda: c5_slab(512)]
    for fn in steps:
        if time.perf_counter() - t0 > budget_s and world == 1:
            out.append({"skipped": "secondary time budget exhausted"})
            break
        try:
            out += fn()
        except Exception as e:  # a secondary measurement must never take the headline line down (a failure that is
            # the same on every rank — an API error, an unsupported size — is skipped on every rank alike)
            out.append({"error": f"{type(e).__name__}: {e}"})
            torch.cuda.empty_cache()
    return out
