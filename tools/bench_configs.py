"""Secondary BASELINE configs (3, 4, 5) on one GPU, or config 5 slab-decomposed under torchrun:
one JSON line per measurement.  bench.py (config 2) stays the contract line; these explain the
other rows of SURVEY section 8.  Usage:
    python tools/bench_configs.py [--only c3|c4|c5] [--n3 512]
    torchrun --nproc-per-node P tools/bench_configs.py --only c5slab --n3 512
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, warmup=2, iters=5):
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


def fp32_peak():
    """FFMA-chain peak measured live on this GPU (TFLOP/s): the denominator of the FP32-bound rooflines."""
    import ctypes

    from pde_opt_b200 import _lib

    v = ctypes.c_double(0.0)
    _lib.check(_lib.load().pdeopt_measure_fp32_peak(ctypes.byref(v), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return float(v.value)


def peak_hbm():
    try:
        return json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6538.3


def c1(args):
    """BASELINE config 1: Allen-Cahn 64x64, single env, 1000 semi-implicit steps (latency-bound: one CTA)."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AllenCahn2DPeriodic
    from pde_opt_b200.functions import ConstantMobility, DoubleWell
    from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral

    n = 64
    dom = Domain((n, n), ((0.0, 0.01 * n),) * 2, "dimensionless")
    eq = AllenCahn2DPeriodic(dom, 0.002, DoubleWell(), ConstantMobility(1.0))
    solver = SemiImplicitFourierSpectral(1.0, eq.fourier_symbol, eq.fft, eq.ifft)
    y = torch.from_numpy((0.01 * np.random.default_rng(0).normal(size=(n, n))).astype(np.float32)).cuda()
    times = (np.arange(1001, dtype=np.float64) * 5e-6).astype(np.float32)
    t = timed(lambda: solver.rollout(ODETerm(eq), times, y), 2, 5)
    for B in (1, 296):
        yb = y[None].repeat(B, 1, 1).contiguous()
        tb = timed(lambda: solver.rollout(ODETerm(eq), times, yb), 2, 5)
        print(json.dumps({"config": "C1 Allen-Cahn 64x64, 1000 steps (small-grid kernel)", "envs": B, "ms_per_1000_steps": tb * 1e3,
                          "env_steps_per_s": B * 1000 / tb, "note": "single env = one CTA: latency-bound"}))
    del t


def c3(args):
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import GPE2DTSControl
    from pde_opt_b200.solvers import ODETerm, StrangSplitting

    N, B, K = 256, args.envs3, 16
    L_ = 29.4
    dom = Domain((N, N), ((-L_ / 2, L_ / 2),) * 2, "dimensionless")
    eq = GPE2DTSControl(dom, 3371.7, 0.0, None, 1.0)
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gpe_ground_state_256.npy"))
    rng = np.random.default_rng(0)
    y0 = np.stack([g * (1 + 0.01 * rng.normal(size=g.shape)) for _ in range(B)]).astype(np.float32)
    y = torch.from_numpy(y0).cuda()
    times = np.arange(K + 1, dtype=np.float32) * np.float32(2 * np.pi * 1e-4)
    for kinetic in (True, False):
        a = (0.5j * eq.two_pi_i_k_2).astype(np.complex64) if kinetic else eq.A_term
        for ts in (-1j, 1.0):
            solver = StrangSplitting(a, eq.dx, eq.fft, eq.ifft, ts)
            out = torch.empty_like(y)
            t = timed(lambda: solver.rollout(ODETerm(eq), times, y, out=out))
            flop = (4 * 5 * 16 + 40 if kinetic else 40) * N * N
            passes = 8 if kinetic else 4  # 4 kernels per step, each one read + one write of the state
            print(json.dumps({"config": "C3 GPE 256x256 c64 Strang", "kinetic": kinetic, "time_scale": str(ts), "envs": B, "steps": K,
                              "env_steps_per_s": B * K / t, "ms_per_step": t / K * 1e3,
                              "tflops_algorithmic": flop * B * K / t / 1e12,
                              "kernel_traffic_GBps": passes * N * N * 8 * B * K / t / 1e9, "note": "line-FFT path, 4 kernels per step, state L2-resident (64 MB)"}))


def c3b(args):
    """GPE 128x128 (the reference's own test size, tests/test_solvers.py:107-205) on the fused single-CTA kernel."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import GPE2DTSControl
    from pde_opt_b200.solvers import ODETerm, StrangSplitting

    N, B, K = 128, 2048, 16
    L_ = 29.4
    dom = Domain((N, N), ((-L_ / 2, L_ / 2),) * 2, "dimensionless")
    eq = GPE2DTSControl(dom, 3371.7, 0.0, None, 1.0)
    rng = np.random.default_rng(0)
    psi = rng.normal(size=(B, N, N, 2)).astype(np.float32)
    psi /= np.sqrt((psi**2).sum(axis=(1, 2, 3), keepdims=True) * eq.dx**2)
    y = torch.from_numpy(psi).cuda()
    times = np.arange(K + 1, dtype=np.float32) * np.float32(2 * np.pi * 1e-4)
    for kinetic in (True, False):
        a = (0.5j * eq.two_pi_i_k_2).astype(np.complex64) if kinetic else eq.A_term
        solver = StrangSplitting(a, eq.dx, eq.fft, eq.ifft, -1j)
        out = torch.empty_like(y)
        t = timed(lambda: solver.rollout(ODETerm(eq), times, y, out=out))
        flop = (4 * 5 * 14 + 40 if kinetic else 40) * N * N
        print(json.dumps({"config": "GPE 128x128 c64 Strang (fused single-CTA kernel)", "kinetic": kinetic, "envs": B, "steps": K,
                          "env_steps_per_s": B * K / t, "tflops_algorithmic": flop * B * K / t / 1e12}))


def c4(args):
    from pde_opt_b200 import Domain
    from pde_opt_b200.adjoint import ad_rollout
    from pde_opt_b200.equations import AdvectionDiffusion2D
    from pde_opt_b200.functions import GaussianVelocity

    N, H, B, K = 128, 0.02, args.envs4, 500
    dom = Domain((N, N), ((-N * H / 2, N * H / 2),) * 2, "dimensionless")
    eq = AdvectionDiffusion2D(dom, GaussianVelocity(0.1, 0.01), 0.1)
    rng = np.random.default_rng(0)
    y0 = torch.from_numpy((0.5 + 0.01 * rng.normal(size=(B, N, N))).astype(np.float32)).cuda()
    ctrl = torch.tensor([0.1, -0.1, 0.1, 0.01], device="cuda").expand(B, 10, 4).contiguous()
    times = np.arange(K + 1, dtype=np.float32) * np.float32(1e-4)
    t_f = timed(lambda: ad_rollout(eq, y0, ctrl, times, hold=50), 1, 3)

    def fb():
        yg, cg = y0.clone().requires_grad_(True), ctrl.clone().requires_grad_(True)
        (ad_rollout(eq, yg, cg, times, hold=50) ** 2).mean().backward()

    t_fb = timed(fb, 1, 3)
    flop = 2.5e6
    print(json.dumps({"config": "C4 advection-diffusion 128x128, 500-step rollout", "envs": B, "steps": K,
                      "forward_env_steps_per_s": B * K / t_f, "forward_tflops": flop * B * K / t_f / 1e12,
                      "forward_plus_adjoint_env_steps_per_s": B * K / t_fb, "fwd_bwd_tflops": 2 * flop * B * K / t_fb / 1e12,
                      "trajectory_GiB": B * K * N * N * 4 / 2**30}))


def c5(args):
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard3DPeriodic
    from pde_opt_b200.functions import ConstantMobility, LogRegular
    from pde_opt_b200.linefft import pos_to_freq

    n = args.n3
    pts = (n, n, n)
    dom = Domain(pts, tuple((0.0, n * 0.01) for _ in range(3)), "dimensionless")
    eq = CahnHilliard3DPeriodic(dom, 0.002, LogRegular(3.0), ConstantMobility(0.15))
    # position-ordered symbol on the device without the 1 GB host array: separable |k|^2
    kn = (2 * np.pi * np.fft.fftfreq(n, 0.01)).astype(np.float32)
    k = [torch.as_tensor(kn[pos_to_freq(n)] ** 2, device="cuda") for _ in range(2)] + [torch.as_tensor(kn[: n // 2 + 1] ** 2, device="cuda")]
    k2 = (k[0][:, None, None] + k[1][None, :, None]) + k[2][None, None, :]
    sym = (0.5 * 0.002 * k2 * k2).contiguous()
    del k2
    u = torch.from_numpy(np.clip(0.5 + 0.01 * np.random.default_rng(0).normal(size=(1,) + pts), 0.01, 0.99).astype(np.float32)).cuda()
    out = torch.empty_like(u)
    plan = eq.plan()
    K = 4
    dts = np.full(K, 1e-6, np.float32)
    t = timed(lambda: plan.step(u, dts, sym, out=out), 1, 3) / K
    vol = n**3
    alg = (2 * 4 * vol) + 2 * 3 * 2 * (8 * (n // 2 + 1) * n * n)
    print(json.dumps({"config": f"C5 Cahn-Hilliard 3D {n}^3 single GPU", "ms_per_step": t * 1e3, "steps_per_s": 1 / t,
                      "grid_point_steps_per_s": vol / t, "algorithmic_GB_per_step": alg / 1e9, "achieved_GBps": alg / t / 1e9,
                      "hbm_peak_GBps": peak_hbm(), "frac": alg / t / 1e9 / peak_hbm(), "finite": bool(torch.isfinite(out).all())}))


def c5slab(args):
    import torch.distributed as dist

    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard3DPeriodic
    from pde_opt_b200.functions import ConstantMobility, LogRegular
    from pde_opt_b200.linefft import pos_to_freq
    from pde_opt_b200.parallel import SlabCahnHilliard3D

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    n = args.n3
    pts = (n, n, n)
    dom = Domain(pts, tuple((0.0, n * 0.01) for _ in range(3)), "dimensionless")
    eq = CahnHilliard3DPeriodic(dom, 0.002, LogRegular(3.0), ConstantMobility(0.15))
    C, nxl = n // world, n // world
    pf = pos_to_freq(n)
    knat = (2 * np.pi * np.fft.fftfreq(n, 0.01)).astype(np.float32)
    kk = knat[pf] ** 2
    kx = torch.as_tensor(kk, device="cuda")
    ky = torch.as_tensor(kk[rank * C : (rank + 1) * C], device="cuda")
    kz = torch.as_tensor(knat[: n // 2 + 1] ** 2, device="cuda")
    k2 = (kx[:, None, None] + ky[None, :, None]) + kz[None, None, :]
    sym = (0.5 * 0.002 * k2 * k2).contiguous()
    del k2
    slab = SlabCahnHilliard3D(eq, 0.5, device=torch.device("cuda", torch.cuda.current_device()), symbol_pos_local=sym, transport=args.transport)
    full = np.clip(0.5 + 0.01 * np.random.default_rng(0).normal(size=pts), 0.01, 0.99).astype(np.float32)
    u = torch.from_numpy(full[rank * nxl : (rank + 1) * nxl].copy()).cuda()
    out = torch.empty_like(u)
    for _ in range(2):
        slab.step(u, 1e-6, out=out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if slab._timing is not None:
        slab._timing = []  # drop the warm-up steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    e0.record()
    for _ in range(iters):
        slab.step(u, 1e-6, out=out)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e-3], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if slab._timing is not None and rank == 0:
        print(json.dumps({"phase_ms_rank0": {k: round(v, 4) for k, v in slab.phase_times_ms().items()}}))
    chk = out.double().sum()
    if world > 1:
        dist.all_reduce(chk)
    parity = None
    if args.check:
        # gather the slabs and compare with the whole-domain single-GPU step on rank 0
        allo = torch.empty((world * nxl, n, n), dtype=torch.float32, device="cuda")
        if world > 1:
            dist.all_gather_into_tensor(allo, out)
        else:
            allo.copy_(out)
        if rank == 0:
            kfull = torch.as_tensor(kk, device="cuda")
            k2f = (kfull[:, None, None] + kfull[None, :, None]) + kz[None, None, :]
            symf = (0.5 * 0.002 * k2f * k2f).contiguous()
            ref = eq.plan().step(torch.from_numpy(full[None]).cuda(), np.asarray([1e-6], np.float32), symf)[0]
            u0 = torch.from_numpy(full).cuda()
            parity = {"rel_l2_state": float(((allo - ref).norm() / ref.norm()).item()),
                      "rel_l2_increment": float((((allo - u0) - (ref - u0)).norm() / (ref - u0).norm()).item())}
    if rank == 0:
        tt = float(t.item())
        print(json.dumps({"config": f"C5 Cahn-Hilliard 3D {n}^3 slab-decomposed", "transport": slab.transport, "n_gpus": world, "ms_per_step": tt * 1e3,
                          "grid_point_steps_per_s": n**3 / tt, "all_to_all_MB_per_rank_per_step": 2 * nxl * n * n * 8 * (world - 1) / world / 1e6,
                          "checksum": float(chk.item()), "parity_vs_single_gpu": parity}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--n3", type=int, default=512)
    ap.add_argument("--envs3", type=int, default=128)
    ap.add_argument("--envs4", type=int, default=512)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--transport", default="auto", choices=["auto", "nccl", "peer", "push"])
    a = ap.parse_args()
    todo = [a.only] if a.only else ["c1", "c3", "c3b", "c4", "c5"]
    for name in todo:
        {"c1": c1, "c3": c3, "c3b": c3b, "c4": c4, "c5": c5, "c5slab": c5slab}[name](a)
