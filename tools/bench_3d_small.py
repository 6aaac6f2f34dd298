"""Small 3-D Cahn-Hilliard grids (docs/notebooks/optimization_3D.ipynb sizes): steps/s of pdeopt_ch3d_step."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pde_opt_b200 import Domain
from pde_opt_b200.equations import CahnHilliard3DPeriodic
from pde_opt_b200.functions import ConstantMobility, LogRegular
from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral

for n, B, K in [(32, 1, 64), (64, 1, 64), (64, 8, 64), (128, 1, 32)]:
    pts = (n, n, n)
    dom = Domain(pts, tuple((0.0, n * 0.01) for _ in range(3)), "dimensionless")
    eq = CahnHilliard3DPeriodic(dom, 0.002, LogRegular(3.0), ConstantMobility(0.15))
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    u = torch.from_numpy(np.clip(0.5 + 0.01 * np.random.default_rng(0).normal(size=(B,) + pts), 0.01, 0.99).astype(np.float32)).cuda()
    times = np.arange(K + 1, dtype=np.float32) * np.float32(1e-6)
    for _ in range(3):
        out = solver.rollout(ODETerm(eq), times, u)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        out = solver.rollout(ODETerm(eq), times, u)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / reps / K
    print(json.dumps({"grid": n, "batch": B, "us_per_step": t * 1e6, "grid_point_steps_per_s": B * n**3 / t, "finite": bool(torch.isfinite(out).all())}))
