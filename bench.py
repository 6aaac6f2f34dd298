#!/usr/bin/env python
"""Benchmark of the fused time-stepping hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Cahn-Hilliard 2-D 128x128 float32, 4096
batched PDEEnv instances PER GPU (weak scaling), control forcing on (per-env interaction
offset + Gaussian bump in mu), 16 fused numeric steps per env step, uint8 observation and
(mean, variance) reward computed in the kernel epilogue.  One bench "step" = one env step of
the whole batch = one launch.  Metric: env-steps/s = envs * numeric steps / time.

`value`   : device-resident state, CUDA events, max over ranks.
`e2e`     : same work through pdeopt_sifs_step_batched_host — pinned HOST state/control in,
            state/observation/reward out, copies inside the timed region.
`--impl reference`: the reference's own CPU arithmetic.  jax/diffrax are not installable in
            this image, so this times the NumPy restatement in oracle/ on all host cores
            (cpu_baseline.kind = "port"), on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 128
H = 0.01
KAPPA = 0.002
A_SPLIT = 0.5
DT = 1e-6
K_FUSED = 16
ENVS_PER_GPU = 4096
CPU_SAMPLE_ENVS_PER_CORE = 768  # ~10 s of CPU work per core for the cpu_baseline leg
OMEGA = 3.0
FLOP_PER_POINT = 70 + 35 + 1 + 2  # BASELINE.md section 3, log potential: 108 flop / grid point / numeric step
BYTES_PER_ENV_LAUNCH = 2 * 4 * N * N  # one read + one write of the state per launch (K fused steps)
KERNEL_NAME = "sifs128r_kernel"
# roofline.traffic = dram__bytes_read.sum + dram__bytes_write.sum of one launch of the headline kernel at this exact
# config, read from the committed summary of the `ncu --set full` capture (written by tools/ncu_summary.py --json);
# null when the summary is missing or belongs to another kernel
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_bench_traffic.json")


def ncu_traffic():
    try:
        t = json.load(open(TRAFFIC_FILE))
        if KERNEL_NAME not in t.get("kernel", "") or t.get("grid_envs") != ENVS_PER_GPU or t.get("fused_steps") != K_FUSED:
            return None, None
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"]), os.path.relpath(TRAFFIC_FILE, ROOT) + " <- " + t.get("source", "?")
    except Exception:
        return None, None


def make_ic(env_index):
    rng = np.random.default_rng(env_index)
    return np.clip(0.5 + 0.01 * rng.normal(size=(N, N)), 0.0, 1.0).astype(np.float32)


def make_ctrl(env_indices):
    c = np.zeros((len(env_indices), 8), np.float32)
    for i, e in enumerate(env_indices):
        rng = np.random.default_rng(10_000_000 + e)
        c[i, 0] = rng.uniform(-0.2, 0.2)          # interaction offset
        c[i, 1] = rng.uniform(-0.5, 0.5)          # bump amplitude
        c[i, 2:4] = rng.uniform(-0.4, 0.4, 2)     # bump centre
        c[i, 4] = rng.uniform(0.05, 0.2)          # bump width
    return c


def symbol_quadrant():
    k = np.fft.fftfreq(N, H)[: N // 2 + 1]
    two_pi_i_k2 = -((2 * np.pi * k[:, None]) ** 2 + (2 * np.pi * k[None, :]) ** 2)
    return (KAPPA * two_pi_i_k2**2).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference arithmetic)
# ----------------------------------------------------------------------------------------------

def _cpu_worker(args):
    env_ids, nsteps = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import pde_oracle as O

    dom = O.Domain((N, N), ((-N * H / 2, N * H / 2),) * 2)
    X, Y = dom.mesh(np.float32)
    ctrl = make_ctrl(env_ids)
    t0 = time.perf_counter()
    acc = 0.0
    for i, e in enumerate(env_ids):
        w = OMEGA + float(ctrl[i, 0])
        bump = (ctrl[i, 1] * np.exp(-((X - ctrl[i, 2]) ** 2 + (Y - ctrl[i, 3]) ** 2) / (2 * ctrl[i, 4] ** 2))).astype(np.float32)
        eq = O.CahnHilliardPeriodic(dom, KAPPA, lambda c, w=w: O.mu_log(c, w), lambda c: (1 - c) * c, "fd", np.float32, forcing=bump)
        y = make_ic(e)
        t = np.float32(0)
        for _ in range(nsteps):
            y = O.sifs_step(eq.rhs, y, t, t + np.float32(DT), A_SPLIT, eq.fourier_symbol)
            t = t + np.float32(DT)
        acc += float(y.var())
    return time.perf_counter() - t0, acc


def cpu_arm(envs_per_core, nsteps, cores=None):
    """env-steps/s of the oracle port over `cores` worker processes."""
    import multiprocessing as mp

    cores = cores or (os.cpu_count() or 1)
    jobs = [(list(range(c * envs_per_core, (c + 1) * envs_per_core)), nsteps) for c in range(cores)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    return cores * envs_per_core * nsteps / wall, cores, wall


def run_reference(args):
    """The reference's CPU arithmetic (NumPy restatement; jax/diffrax cannot be installed here) on all
    host cores: a persistent pool of one worker per core, each bench step = cores x 48 environments
    x 16 numeric steps (about one second), so process start-up is outside the timed region."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    envs_per_core = 48
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def one_step(i):
            jobs = [(list(range((i * cores + c) * envs_per_core, (i * cores + c + 1) * envs_per_core)), K_FUSED) for c in range(cores)]
            pool.map(_cpu_worker, jobs)

        for i in range(max(args.warmup, 1)):
            one_step(i)
        t0 = time.perf_counter()
        for i in range(args.steps):
            one_step(1000 + i)
        wall = time.perf_counter() - t0
    value = args.steps * cores * envs_per_core * K_FUSED / wall
    sample = f"{cores * envs_per_core} envs x {K_FUSED} numeric steps per bench step, {args.steps} steps, persistent pool of {cores} workers"
    line = {
        "impl": "reference",
        "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "jax/diffrax not installable here: NumPy restatement of the reference arithmetic (oracle/), one process per host core",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------

def workload_config(n_gpus):
    return {
        "workload": "Cahn-Hilliard 2D 128x128 fp32, rhs_fd, mu=log(c/(1-c))+w(1-2c), D=c(1-c), A=0.5, dt=1e-6, "
                    "control forcing on, K=16 fused steps per env step, uint8 obs + (mean,var) reward epilogue",
        "envs_per_gpu": ENVS_PER_GPU, "global_envs": ENVS_PER_GPU * n_gpus, "fused_steps": K_FUSED, "grid": [N, N],
        "parallelism": f"env-sharded x{n_gpus}, no collective in step, all_gather of rewards",
        "l2_policy": "state per GPU is 256 MiB in + 256 MiB out (> 126 MB L2); ping-pong buffers",
        "initial_conditions": "clip(0.5 + 0.01 N(0,1), 0, 1) drawn on the device with torch.Generator(seed = 1234 + rank) "
                              "(timing only; the parity tests use numpy default_rng(env_index), SURVEY 8d)",
    }


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML, ~20 ms
    period; falls back to nvidia-smi)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.samples, self.reasons = gpu_index, [], set()
        self._halt = threading.Event()
        self.max_mhz = None

    def _nvml_loop(self):
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._halt.is_set():
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for nm, b in bits.items():
                if r & b:
                    self.reasons.add(nm)
            self._halt.wait(0.02)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.1)

    def run(self):
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from pde_opt_b200 import _lib
    from pde_opt_b200.fused import SifsPlan
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the stepping path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B = ENVS_PER_GPU
    env0 = rank * B
    plan = SifsPlan("ch2d", N, N, (-N * H / 2,) * 2, (H, H), KAPPA, ("log", (OMEGA,)), ("degenerate", ()))
    sym_host = np.ascontiguousarray(np.float32(A_SPLIT) * symbol_quadrant())
    sym = torch.from_numpy(sym_host).to(dev)
    # synthetic initial conditions: distribution of notebooks/optimize_nn_script.py:40 (clip(0.5+0.01 N, 0, 1));
    # generated on the device for speed (seeded per rank), the parity tests use numpy default_rng(env_index).
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    y_a = (0.5 + 0.01 * torch.randn((B, N, N), device=dev, generator=g)).clamp_(0, 1).contiguous()
    y_b = torch.empty_like(y_a)
    ctrl_host = make_ctrl(list(range(env0, env0 + B)))
    ctrl = torch.from_numpy(ctrl_host).to(dev)
    obs = torch.empty((B, N, N), dtype=torch.uint8, device=dev)
    rew = torch.empty((B, 2), dtype=torch.float32, device=dev)
    rew_all = torch.empty((world * B, 2), dtype=torch.float32, device=dev) if world > 1 else None
    dts = [DT] * K_FUSED

    def env_step(src, dst):
        plan.step(src, dts, sym, ctrl=ctrl, obs=obs, obs_range=(0.0, 1.0), reward=rew, out=dst)
        if world > 1:
            dist.all_gather_into_tensor(rew_all, rew)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cur, nxt = y_a, y_b
    for _ in range(max(args.warmup, 3)):
        env_step(cur, nxt)
        cur, nxt = nxt, cur
    barrier()
    launches0 = lib.pdeopt_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        env_step(cur, nxt)
        cur, nxt = nxt, cur
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.pdeopt_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    finite = bool(torch.isfinite(cur).all())
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * K_FUSED / (ms_per_step * 1e-3)

    # ---- e2e: pinned host buffers through the C-ABI host entry point ----
    yh = torch.empty((B, N, N), dtype=torch.float32).pin_memory()
    yh.copy_(cur.cpu())
    yo = torch.empty_like(yh).pin_memory()
    obs_h = torch.empty((B, N, N), dtype=torch.uint8).pin_memory()
    rew_h = torch.empty((B, 2), dtype=torch.float32).pin_memory()
    ctrl_h = torch.from_numpy(ctrl_host).pin_memory()
    e2e_steps = max(2, min(args.steps, 10))

    def e2e_step(a, b):
        plan.step_host(a.numpy(), dts, sym_host, ctrl=ctrl_h.numpy(), obs=obs_h.numpy(), obs_range=(0.0, 1.0),
                       reward=rew_h.numpy(), out=b.numpy())

    for _ in range(2):
        e2e_step(yh, yo)
    barrier()
    t0 = time.perf_counter()
    a, b = yh, yo
    for _ in range(e2e_steps):
        e2e_step(a, b)
        a, b = b, a
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * K_FUSED * e2e_steps / e2e_s

    # ---- copy ceiling of this box: the SAME byte counts as one e2e step, nothing but concurrent H2D and D2H
    # cudaMemcpyAsync calls (torch copy_ on pinned memory; one call per chunk, 8 chunks per direction, two streams),
    # on every rank at once.  e2e can at best equal it: the ratio says how much of the gap to `value` is the box's
    # PCIe / host-memory path and how much is the code's pipelining.
    h2d_b = B * N * N * 4 + B * 8 * 4 + sym_host.nbytes
    d2h_b = B * N * N * 4 + B * N * N + B * 2 * 4
    hin, hout = torch.empty(h2d_b, dtype=torch.uint8).pin_memory(), torch.empty(d2h_b, dtype=torch.uint8).pin_memory()
    din, dout = torch.empty(h2d_b, dtype=torch.uint8, device=dev), torch.empty(d2h_b, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copy_step():
        for c in range(8):
            a0, a1 = c * h2d_b // 8, (c + 1) * h2d_b // 8
            b0, b1 = c * d2h_b // 8, (c + 1) * d2h_b // 8
            with torch.cuda.stream(s_in):
                din[a0:a1].copy_(hin[a0:a1], non_blocking=True)
            with torch.cuda.stream(s_out):
                hout[b0:b1].copy_(dout[b0:b1], non_blocking=True)

    copy_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copy_step()
    torch.cuda.synchronize()
    copy_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([copy_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        copy_s = float(t.item())
    copy_ceiling = world * B * K_FUSED * e2e_steps / copy_s
    del hin, hout, din, dout

    # ---- env-API view: PDEEnv keeps its state on the device (pde_env.py:234-242, 305); per env step
    # the actions (control block) go host->device and the observation + reward come back ----
    # through the public class: PDEVecEnv.step(actions, obs_host, stats_host) steps the batch in slices on
    # separate streams so that the observation copy of one slice overlaps the next slice's kernel
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard2DPeriodic
    from pde_opt_b200.functions import DegenerateMobility, LogRegular
    from pde_opt_b200.pde_env import PDEVecEnv
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    dom = Domain((N, N), ((-N * H / 2, N * H / 2),) * 2, "dimensionless")
    eq = CahnHilliard2DPeriodic(dom, KAPPA, LogRegular(OMEGA), DegenerateMobility())
    env = PDEVecEnv(eq, SemiImplicitFourierSpectral(A_SPLIT, eq.fourier_symbol, eq.fft, eq.ifft), B, end_time=1e9,
                    step_dt=K_FUSED * DT, numeric_dt=DT, reset_func=None,
                    action_to_control=lambda actions, block: block.copy_(actions, non_blocking=True), auto_reset=False)
    env.state.copy_(cur)
    obs_h4 = obs_h.view(B, 1, N, N)
    env_steps = max(3, min(args.steps, 20))

    def env_api_step():
        env.step(ctrl_h, obs_host=obs_h4, stats_host=rew_h)

    env_api_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(env_steps):
        env_api_step()
    env_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([env_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        env_s = float(t.item())
    env_api_value = world * B * K_FUSED * env_steps / env_s
    h2d = B * N * N * 4 + B * 8 * 4 + sym_host.nbytes
    d2h = B * N * N * 4 + B * N * N + B * 2 * 4

    # ---- roofline of the dominant (only) kernel ----
    peak_tf = ctypes.c_double(0.0)
    _lib.check(lib.pdeopt_measure_fp32_peak(ctypes.byref(peak_tf), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    flop_per_launch = FLOP_PER_POINT * N * N * B * K_FUSED
    achieved_tf = flop_per_launch / (ms_per_step * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_achieved = BYTES_PER_ENV_LAUNCH * B / (ms_per_step * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic()

    # ---- secondary configs (every rank takes part under torchrun; bounded) ----
    secondary = None
    if not args.no_secondary:
        torch.cuda.empty_cache()
        from pde_opt_b200 import secondary_bench

        secondary = secondary_bench.run_all(float(peak_tf.value), world)

    if rank == 0:
        cpu_value, cores, cpu_wall = (None, None, None)
        if world == 1 and not args.no_cpu_baseline:
            cpu_value, cores, cpu_wall = cpu_arm(envs_per_core=CPU_SAMPLE_ENVS_PER_CORE, nsteps=K_FUSED)
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "grid_point_steps_per_s": value * N * N,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "pdeopt_sifs_step_batched_host (pinned host state+control in, state+obs+reward out; "
                           "chunks pipelined over 3 streams)",
                    "steps": e2e_steps,
                    "copy_ceiling": {"value": copy_ceiling, "unit": "env-steps/s", "GBs_per_gpu_both_directions": (h2d_b + d2h_b) * e2e_steps / copy_s / 1e9,
                                     "what": "the same bytes per step as plain concurrent H2D + D2H cudaMemcpyAsync from / to pinned memory on all ranks at once, no kernels"},
                    "frac_of_copy_ceiling": e2e_value / copy_ceiling,
                    "env_api": {"value": env_api_value, "unit": "env-steps/s", "h2d_bytes_per_step": B * 8 * 4,
                                "d2h_bytes_per_step": B * N * N + B * 2 * 4, "steps": env_steps,
                                "what": "PDEVecEnv.step(actions, obs_host, stats_host): state resident on the device as in PDEEnv (pde_env.py:305); pinned actions in, uint8 observation + reward out on the host before the call returns (4 slices on 4 streams overlap copy and compute)"}},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp32", "achieved": achieved_tf, "peak": float(peak_tf.value), "unit": "TFLOP/s",
                "frac": achieved_tf / float(peak_tf.value) if peak_tf.value else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "pdeopt::rf::sifs128r_kernel<CH, MU_LOG, MOB_DEGENERATE> (one environment per 256-thread CTA, two CTAs per SM)",
                "how": f"{FLOP_PER_POINT} algorithmic flop/grid-point/step (BASELINE.md s3) x 16384 points x {B} envs x {K_FUSED} steps per launch / "
                       "CUDA-event launch time; peak = FFMA-chain peak measured live on this GPU (pdeopt_measure_fp32_peak); "
                       "MEASURED_PEAKS.json has no FP32 entry",
                "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                        "how": "2*4*128*128 B per env per launch (state read + write once per K fused steps); of measured copy bandwidth"},
            },
            "finite": finite,
        }
        if secondary is not None:
            line["secondary"] = secondary
        if cpu_value is not None:
            line["cpu_baseline"] = {"value": cpu_value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{cores * CPU_SAMPLE_ENVS_PER_CORE} envs x {K_FUSED} numeric steps, NumPy restatement of the reference (oracle/), {cpu_wall:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the bounded C3/C4/C5 measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
