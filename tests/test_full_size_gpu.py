"""BASELINE.json's full sizes on the GPU, checked through size-independent properties.

The oracle cannot run these sizes in seconds, so each configuration is checked by what the domain
guarantees at any size: conservation laws of the discrete operators (the zero mode of the semi-implicit
filter is exactly 1, the FD divergence telescopes), the normalisation the Strang step enforces
(solvers.py:116-120), independence of the environments of a batch (an environment's result must not depend
on who shares its launch: bit-exact against the same environments stepped in a small batch, which in turn
is what the small-size oracle tests pin), and oracle spot checks of single environments taken out of
the full batch.
"""
import numpy as np
import pytest

from oracle import pde_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))


def test_config2_cahn_hilliard_4096_envs_16_steps():
    """C2: 4096 envs x 128^2, K = 16, log potential, D = c(1-c), A = 0.5, dt = 1e-6, per-env control."""
    from pde_opt_b200.fused import SifsPlan, fold_symbol

    N, H, KAPPA, B, K = 128, 0.01, 0.002, 4096, 16
    g = torch.Generator(device="cuda").manual_seed(0)
    y0 = (0.5 + 0.01 * torch.randn((B, N, N), device="cuda", generator=g)).clamp(0.0, 1.0).contiguous()
    ctrl = torch.zeros((B, 8), device="cuda")
    ctrl[:, 0] = torch.linspace(-0.2, 0.2, B, device="cuda")  # per-env offset of the interaction coefficient
    ctrl[:, 4] = 1.0
    dom = O.Domain((N, N), ((-N * H / 2, N * H / 2),) * 2)
    oeq0 = O.CahnHilliardPeriodic(dom, KAPPA, lambda c: O.mu_log(c, 3.0), lambda c: (1 - c) * c, "fd", np.float32)
    sym = torch.from_numpy(fold_symbol(oeq0.fourier_symbol, 0.5)).cuda()
    plan = SifsPlan("ch2d", N, N, (-N * H / 2,) * 2, (H, H), KAPPA, ("log", (3.0,)), ("degenerate", ()))
    # the float32 time grid of the constant-step loop (dt0 accumulated in float32, last step clipped to t1)
    times = O.constant_step_schedule(0.0, K * 1e-6, 1e-6, np.float32)
    dts = [float(d) for d in (times[1:] - times[:-1])]
    y1 = plan.step(y0, dts, sym, ctrl=ctrl)
    assert bool(torch.isfinite(y1).all())
    # mass conservation per environment: div of face fluxes sums to zero, filter zero mode is 1 / (1 + 0)
    m0, m1 = y0.double().mean(dim=(1, 2)), y1.double().mean(dim=(1, 2))
    assert float((m1 - m0).abs().max()) <= 2e-7
    # every environment actually moved, none blew up
    inc = (y1 - y0).double().flatten(1).norm(dim=1)
    assert float(inc.min()) > 0 and float(inc.max()) < 10 * float(inc.median())
    # independence of the batch: the same pairs stepped alone give bit-identical results
    for lo in (0, 1234, 4094):
        sub = plan.step(y0[lo:lo + 2].contiguous(), dts, sym, ctrl=ctrl[lo:lo + 2].contiguous())
        assert torch.equal(sub, y1[lo:lo + 2])
    # oracle spot check of the first and the last environment
    for b in (0, B - 1):
        w = 3.0 + float(ctrl[b, 0])
        oeq = O.CahnHilliardPeriodic(dom, KAPPA, lambda c, w=w: O.mu_log(c, w), lambda c: (1 - c) * c, "fd", np.float32)
        y = y0[b].cpu().numpy()
        for a, bb in zip(times[:-1], times[1:]):
            y = O.sifs_step(oeq.rhs, y, a, bb, 0.5, oeq.fourier_symbol)
        assert rel_l2(y1[b].cpu().numpy(), y) <= 1e-5
        assert rel_l2(y1[b].cpu().numpy() - y0[b].cpu().numpy(), y - y0[b].cpu().numpy()) <= 2e-3


@pytest.mark.parametrize("time_scale", [-1j, 1.0])
def test_config3_gpe_256_128_envs(time_scale):
    """C3 per-GPU share: 128 envs x 256^2 complex64, kinetic term on, reference test parameters."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import GPE2DTSControl
    from pde_opt_b200.solvers import ODETerm, StrangSplitting
    from tests.test_strang_gpu import tf_setup

    L_, k, x_s, t_s = tf_setup()
    n, B, K = 256, 128, 8
    box = ((-L_ / 2, L_ / 2), (-L_ / 2, L_ / 2))
    dom, odom = Domain((n, n), box, "dimensionless"), O.Domain((n, n), box)
    eq = GPE2DTSControl(dom, k, 0.0, lambda t, x, y: 0.0 * x, trap_factor=1.0)
    oeq = O.GPE2DTSControl(odom, k, 0.0, lambda t, x, y: 0.0 * x, 1.0, np.float32, kinetic=True)
    solver = StrangSplitting(oeq.A_term, eq.dx, eq.fft, eq.ifft, time_scale)
    dx2 = float(odom.dx[0]) ** 2
    g = torch.Generator(device="cuda").manual_seed(1)
    i = torch.arange(n, device="cuda", dtype=torch.float32)
    env = torch.exp(-(((i[:, None] - n / 2) / (0.3 * n)) ** 2) - ((i[None, :] - n / 2) / (0.3 * n)) ** 2)
    psi = env[None, :, :, None] * (1.0 + 0.05 * torch.randn((B, n, n, 2), device="cuda", generator=g))
    psi = psi / torch.sqrt((psi.double() ** 2).sum(dim=(1, 2, 3), keepdim=True) * dx2).float()
    dt_ = 1e-5 / t_s
    times = O.constant_step_schedule(0.0, K * dt_, dt_, np.float32)
    out = solver.rollout(ODETerm(eq), times, psi.contiguous())
    assert bool(torch.isfinite(out).all())
    # solvers.py:116-120 renormalises to sum |psi|^2 dx^2 = 1 after the potential step; the closing
    # half step of the kinetic term follows it: unitary in real time (norm stays 1 to rounding), a slight
    # decay of O(dt E_kin) in imaginary time
    norm = (out.double() ** 2).sum(dim=(1, 2, 3)) * dx2
    assert float((norm - 1.0).abs().max()) <= (2e-5 if time_scale == 1.0 else 1e-3)
    # independence of the batch (the per-environment norm is accumulated with float atomics, so the last
    # bits depend on the order of the tiles: tolerance, not equality) and an oracle spot check
    for b in (0, 77, B - 1):
        alone = solver.rollout(ODETerm(eq), times, psi[b:b + 1].contiguous())
        assert rel_l2(alone[0].cpu().numpy(), out[b].cpu().numpy()) <= 1e-6
    y = psi[5].cpu().numpy()
    for a, bb in zip(times[:-1], times[1:]):
        y = O.strang_step(oeq.B_terms, y, a, bb, oeq.A_term, oeq.dx, time_scale)
    assert rel_l2(out[5].cpu().numpy(), y) <= 2e-5
    assert abs(float(norm[5]) - float(np.sum(y.astype(np.float64) ** 2) * dx2)) <= 2e-6


def test_config4_advection_diffusion_512_envs_500_steps():
    """C4: 512 envs x 128^2 x 500 steps, Gaussian velocity with a moving centre per env."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.adjoint import ad_rollout
    from pde_opt_b200.equations import AdvectionDiffusion2D
    from pde_opt_b200.functions import GaussianVelocity

    N, H, B, K, nseg = 128, 0.02, 512, 500, 10
    box = ((-N * H / 2, N * H / 2),) * 2
    eq = AdvectionDiffusion2D(Domain((N, N), box, "dimensionless"), GaussianVelocity(0.1, 0.01), 0.1)
    g = torch.Generator(device="cuda").manual_seed(2)
    y0 = (0.5 + 0.01 * torch.randn((B, N, N), device="cuda", generator=g)).contiguous()
    rng = np.random.default_rng(4)
    ctrl = np.empty((B, nseg, 4), np.float32)
    ctrl[..., 0:2] = rng.uniform(-0.5, 0.5, (B, nseg, 2))
    ctrl[..., 2] = rng.uniform(0.05, 0.2, (B, nseg))
    ctrl[..., 3] = rng.uniform(0.01, 0.05, (B, nseg))
    ctrl_d = torch.from_numpy(ctrl).cuda()
    times = (np.arange(K + 1, dtype=np.float32) * np.float32(1e-4)).astype(np.float32)
    out = ad_rollout(eq, y0, ctrl_d, times, hold=K // nseg)
    assert bool(torch.isfinite(out).all())
    # conservative form: -div(v u) + D lap(u) has zero mean, and the filter's zero mode is 1
    m0, m1 = y0.double().mean(dim=(1, 2)), out.double().mean(dim=(1, 2))
    assert float((m1 - m0).abs().max()) <= 5e-6
    assert float(out.abs().max()) < 1e3
    for lo in (0, 300, B - 2):
        sub = ad_rollout(eq, y0[lo:lo + 2].contiguous(), ctrl_d[lo:lo + 2].contiguous(), times, hold=K // nseg)
        assert torch.equal(sub, out[lo:lo + 2])


def test_config5_cahn_hilliard_512_cubed_one_step_properties():
    """C5: one 512^3 domain (dx = 0.01, kappa = 0.002, log potential, D = 0.15, A = 0.5, dt = 1e-6)."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard3DPeriodic
    from pde_opt_b200.functions import ConstantMobility, LogRegular
    from pde_opt_b200.linefft import pos_to_freq

    n = 512
    pts = (n, n, n)
    dom = Domain(pts, tuple((0.0, n * 0.01) for _ in range(3)), "dimensionless")
    eq = CahnHilliard3DPeriodic(dom, 0.002, LogRegular(3.0), ConstantMobility(0.15))
    kn = (2 * np.pi * np.fft.fftfreq(n, 0.01)).astype(np.float32)
    k = [torch.as_tensor(kn[pos_to_freq(n)] ** 2, device="cuda") for _ in range(2)] + [torch.as_tensor(kn[: n // 2 + 1] ** 2, device="cuda")]
    k2 = (k[0][:, None, None] + k[1][None, :, None]) + k[2][None, None, :]
    sym = (0.5 * 0.002 * k2 * k2).contiguous()
    del k2
    g = torch.Generator(device="cuda").manual_seed(3)
    u = (0.5 + 0.01 * torch.randn((1,) + pts, device="cuda", generator=g)).clamp(0.01, 0.99).contiguous()
    plan = eq.plan()
    dts = np.full(1, 1e-6, np.float32)
    y1 = plan.step(u, dts, sym)
    assert bool(torch.isfinite(y1).all())
    # mass conservation: the FD divergence telescopes over the periodic grid and the zero mode passes unchanged
    assert abs(float(y1.double().mean()) - float(u.double().mean())) <= 1e-7
    inc = y1 - u
    assert float(inc.abs().max()) > 0
    # the RHS alone: zero mean to rounding, and translation equivariance — a shift by whole tiles of the
    # marching kernel (64 planes, 16 rows, 64 columns) must reproduce the same bits in shifted places
    f = eq.rhs(u[0])
    assert abs(float(f.double().mean())) <= 1e-6 * float(f.double().abs().mean())
    shift = (64, 16, 64)
    f_s = eq.rhs(torch.roll(u[0], shift, dims=(0, 1, 2)).contiguous())
    assert torch.equal(f_s, torch.roll(f, shift, dims=(0, 1, 2)))
    del f, f_s
    # translation equivariance of the whole step (FFT summation order changes with the shift: tolerance)
    y1_s = plan.step(torch.roll(u, shift, dims=(1, 2, 3)).contiguous(), dts, sym)
    num = float((torch.roll(y1_s, tuple(-s for s in shift), dims=(1, 2, 3)) - y1).double().norm())
    assert num / float(inc.double().norm()) <= 1e-3
    # a spatially constant state is a fixed point (mu constant, all fluxes vanish)
    c = torch.full((1,) + pts, 0.37, device="cuda")
    assert float((plan.step(c, dts, sym) - c).abs().max()) <= 1e-7
