"""One small invocation of every kernel family (for compute-sanitizer runs: memcheck / racecheck).
Prints 'sanity ok' when every result is finite."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pde_opt_b200 import Domain  # noqa: E402
from pde_opt_b200.adjoint import ad_rollout  # noqa: E402
from pde_opt_b200.equations import (AdvectionDiffusion2D, AllenCahn2DPeriodic, CahnHilliard2DPeriodic,  # noqa: E402
                                    CahnHilliard3DPeriodic, GPE2DTSControl)
from pde_opt_b200.functions import (ConstantMobility, DegenerateMobility, GaussianLight, GaussianVelocity,  # noqa: E402
                                    LogRegular)
from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral, StrangSplitting  # noqa: E402

which = sys.argv[1:] or ["sifs", "generic", "fourier", "strang", "ad", "lines3d", "strang_lines"]
rng = np.random.default_rng(0)
ok = True


def u0(shape):
    return torch.from_numpy(np.clip(0.5 + 0.01 * rng.normal(size=shape), 0.01, 0.99).astype(np.float32)).cuda()


def box(points, h=0.01):
    return tuple((-n * h / 2, n * h / 2) for n in points)


times = np.arange(3, dtype=np.float32) * np.float32(1e-6)
if "sifs" in which:
    eq = CahnHilliard2DPeriodic(Domain((128, 128), box((128, 128)), "d"), 0.002, LogRegular(3.0), DegenerateMobility())
    s = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    ok &= bool(torch.isfinite(s.rollout(ODETerm(eq), times, u0((3, 128, 128)))).all())
if "generic" in which:
    eq = AllenCahn2DPeriodic(Domain((64, 32), box((64, 32)), "d"), 0.002, LogRegular(3.0), ConstantMobility(1.0))
    s = SemiImplicitFourierSpectral(1.0, eq.fourier_symbol, eq.fft, eq.ifft)
    ok &= bool(torch.isfinite(s.rollout(ODETerm(eq), times, u0((3, 64, 32)))).all())
if "fourier" in which:
    eq = CahnHilliard2DPeriodic(Domain((128, 128), box((128, 128)), "d"), 0.002, LogRegular(3.0), DegenerateMobility(), derivs="fourier")
    s = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    ok &= bool(torch.isfinite(s.rollout(ODETerm(eq), times, u0((2, 128, 128)))).all())
if "strang" in which or "strang_lines" in which:
    for n in ([128] if "strang" in which else []) + ([64] if "strang_lines" in which else []):
        dom = Domain((n, n), ((-10.0, 10.0),) * 2, "d")
        eq = GPE2DTSControl(dom, 100.0, 0.1, GaussianLight(1.0, 1.0, -1.0, 2.0), 1.0)
        a = (0.5j * eq.two_pi_i_k_2).astype(np.complex64)
        s = StrangSplitting(a, eq.dx, eq.fft, eq.ifft, -1j)
        psi = rng.normal(size=(2, n, n, 2)).astype(np.float32)
        psi /= np.sqrt((psi**2).sum(axis=(1, 2, 3), keepdims=True) * eq.dx**2)
        ok &= bool(torch.isfinite(s.rollout(ODETerm(eq), np.arange(3, dtype=np.float32) * np.float32(1e-4), torch.from_numpy(psi).cuda())).all())
if "ad" in which:
    eq = AdvectionDiffusion2D(Domain((128, 128), box((128, 128), 0.02), "d"), GaussianVelocity(0.1, 0.01), 0.1)
    y = u0((3, 128, 128)).requires_grad_(True)
    c = eq.control_block(3, "cuda", 2).clone().requires_grad_(True)
    out = ad_rollout(eq, y, c, np.arange(5, dtype=np.float32) * np.float32(1e-4), hold=2)
    (out**2).mean().backward()
    ok &= bool(torch.isfinite(y.grad).all() and torch.isfinite(c.grad).all())
    eq64 = AdvectionDiffusion2D(Domain((64, 64), box((64, 64), 0.02), "d"), GaussianVelocity(0.1, 0.01), 0.1)
    ok &= bool(torch.isfinite(ad_rollout(eq64, u0((3, 64, 64)), eq64.control_block(3, "cuda"), np.arange(3, dtype=np.float32) * np.float32(1e-4))).all())
if "lines3d" in which:
    for pts in [(16, 32, 64), (8, 8, 16)]:
        eq = CahnHilliard3DPeriodic(Domain(pts, box(pts), "d"), 0.002, LogRegular(3.0), ConstantMobility(0.15))
        s = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
        ok &= bool(torch.isfinite(s.rollout(ODETerm(eq), times, u0((2,) + pts))).all())
torch.cuda.synchronize()
print("sanity ok" if ok else "sanity FAILED (non-finite)")
