"""GPU parity of the fused Strang split-step kernel (GPE 128x128) against the NumPy oracle,
including the reference's Thomas-Fermi known-answer test run on the GPU in float32."""
import numpy as np
import pytest

from oracle import pde_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N = 128


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def tf_setup():
    """parameters of reference tests/test_solvers.py:107-163"""
    atoms, hb, omega = 5e5, 1.05e-34, 2 * np.pi * 10
    omega_z, mass, a_s = np.sqrt(8) * omega, 3.8175406e-26, 100 * 5.29177210903e-11
    x_s, t_s = np.sqrt(hb / (mass * omega)), 1 / omega
    L_ = 150e-6 / x_s
    k = 4 * np.pi * a_s * atoms * np.sqrt((mass * omega_z) / (2 * np.pi * hb))
    return L_, float(k), x_s, t_s


def make(L_, k, lights=None, e=0.0):
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import GPE2DTSControl

    dom = Domain((N, N), ((-L_ / 2, L_ / 2), (-L_ / 2, L_ / 2)), "dimensionless")
    odom = O.Domain((N, N), ((-L_ / 2, L_ / 2), (-L_ / 2, L_ / 2)))
    eq = GPE2DTSControl(dom, k, e, lights if lights is not None else (lambda t, x, y: 0.0 * x), trap_factor=1.0)
    return dom, odom, eq


def psi_ic(odom, x_s, seed=None):
    psi0 = O.initialize_Psi(N, width=100) * x_s
    if seed is not None:
        rng = np.random.default_rng(seed)
        psi0 = psi0 * np.exp(0.3j * rng.normal(size=(N, N))) * (1 + 0.05 * rng.normal(size=(N, N)))
    psi0 = psi0 / np.sqrt(np.sum(np.abs(psi0) ** 2) * odom.dx[0] ** 2)
    return np.stack([psi0.real, psi0.imag], -1).astype(np.float32)


@pytest.mark.parametrize("time_scale", [-1j, 1.0])
@pytest.mark.parametrize("kinetic", [False, True])
def test_strang_steps_match_oracle(time_scale, kinetic):
    from pde_opt_b200.functions import GaussianLight
    from pde_opt_b200.solvers import ODETerm, StrangSplitting

    L_, k, x_s, t_s = tf_setup()
    light = GaussianLight(5.0, 1.5, -2.0, 3.0)
    dom, odom, eq = make(L_, k, light, e=0.1)
    oeq = O.GPE2DTSControl(odom, k, 0.1, lambda t, x, y: light(t, x, y), 1.0, np.float32, kinetic=kinetic)
    a_term = oeq.A_term if kinetic else eq.A_term
    solver = StrangSplitting(a_term, eq.dx, eq.fft, eq.ifft, time_scale)
    y0 = np.stack([psi_ic(odom, x_s, s) for s in range(3)])
    dt_ = 1e-5 / t_s
    times = O.constant_step_schedule(0.0, 8 * dt_, dt_, np.float32)
    got = solver.rollout(ODETerm(eq), times, torch.from_numpy(y0).cuda()).cpu().numpy()
    for b in range(3):
        y = y0[b]
        for a, bb in zip(times[:-1], times[1:]):
            y = O.strang_step(oeq.B_terms, y, a, bb, oeq.A_term, oeq.dx, time_scale)
        assert rel_l2(got[b], y) <= 2e-5, (time_scale, kinetic, b)
        # norm is pinned to 1 every step (solvers.py:111) unless the last half kinetic step changes it
        nrm = np.sum(got[b].astype(np.float64) ** 2) * odom.dx[0] ** 2
        if not kinetic or time_scale == 1.0:
            np.testing.assert_allclose(nrm, 1.0, rtol=1e-4)
    y1, err, dense, st, res = solver.step(ODETerm(eq), times[0], times[1], torch.from_numpy(y0[0]).cuda())
    assert err is None and st is None and res == 0 and y1.shape == (N, N, 2)
    ref1 = O.strang_step(oeq.B_terms, y0[0], times[0], times[1], oeq.A_term, oeq.dx, time_scale)
    assert rel_l2(y1.cpu().numpy(), ref1) <= 1e-5


def test_reference_kat_thomas_fermi_on_gpu():
    """reference tests/test_solvers.py:293-392 (test_2d_gross_pitaevskii_pde_model) through our
    PDEModel.solve on the GPU in float32: imaginary-time relaxation to the Thomas-Fermi density."""
    from pde_opt_b200.equations import GPE2DTSControl
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.solvers import StrangSplitting

    L_, k, x_s, t_s = tf_setup()
    dom, odom, _ = make(L_, k)
    model = PDEModel(GPE2DTSControl, dom, StrangSplitting)
    y0 = psi_ic(odom, x_s)
    t1_, dt_ = 0.1 / t_s, 1e-5 / t_s
    ys = model.solve({"k": k, "e": 0.0, "lights": lambda t, x, y: 0.0 * x, "trap_factor": 1.0}, torch.from_numpy(y0).cuda(),
                     np.linspace(0.0, t1_, 100), {"time_scale": -1j}, dt0=dt_, max_steps=1000000)
    assert tuple(ys.shape) == (100, N, N, 2)
    X, Y = odom.mesh()
    mu = np.sqrt((1.0 * k * np.sqrt(0.5) * np.sqrt(0.5)) / (2.0 * np.pi))
    V = 0.5 * (0.5 * X**2 + 0.5 * Y**2)
    n = np.clip((mu - V) / k, 0.0, None)
    n = n * (1.0 / (np.sum(n) * odom.dx[0] * odom.dx[1] + 1e-12))
    last = ys[-1].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(n, last[..., 0] ** 2 + last[..., 1] ** 2, rtol=1e-3, atol=1e-3)


def test_detect_vortices_matches_oracle():
    """pde_opt/rl_utils.py:19-84 through pdeopt_gpe_detect_vortices: bit-exact integer windings and counts
    against the NumPy restatement, single state (reference result keys) and a batch, with and without
    the amplitude mask; square and rectangular grids, sizes that are not multiples of the tile."""
    from pde_opt_b200.rl_utils import detect_vortices, vortex_counts

    N = 64
    centres = [(20.5, 20.5, 1), (40.5, 24.5, -1), (30.5, 44.5, 1), (44.5, 44.5, 1)]
    psi = O.vortex_test_field(N, centres).astype(np.complex64)
    for amp in (0.0, 0.15):
        want = O.detect_vortices(psi, amp_thresh=amp)
        got = detect_vortices(torch.from_numpy(psi).cuda(), amp_thresh=amp)
        assert np.array_equal(got["winding"].cpu().numpy(), want["winding"])
        for key in ("num_vortices", "total_topological_charge", "abs_charge_count"):
            assert got[key] == want[key]
        assert np.array_equal(got["positions"].cpu().numpy(), want["positions"])
        assert np.array_equal(got["charges"].cpu().numpy(), want["charges"])
    # batch of rotating-frame-like states: random smooth phases with embedded vortices, rectangular grid
    rng = np.random.default_rng(3)
    B, n0, n1 = 5, 40, 70
    i, j = np.meshgrid(np.arange(n0) + 0.0, np.arange(n1) + 0.0, indexing="ij")
    batch = []
    for b in range(B):
        f = np.exp(-((i - n0 / 2) ** 2 / (0.2 * n0 * n0) + (j - n1 / 2) ** 2 / (0.2 * n1 * n1))).astype(complex)
        for _ in range(b + 1):
            ci, cj = rng.integers(8, n0 - 8) + 0.5, rng.integers(8, n1 - 8) + 0.5
            z = (j - cj) + 1j * (i - ci)
            f *= (z / np.abs(z)) ** int(rng.choice([-1, 1]))
        batch.append(f * np.exp(1j * 0.3 * np.sin(2 * np.pi * i / n0)))
    batch = np.stack(batch).astype(np.complex64)
    counts, w = vortex_counts(torch.view_as_real(torch.from_numpy(batch)).cuda(), amp_thresh=0.1, winding=True)
    for b in range(B):
        want = O.detect_vortices(batch[b], amp_thresh=0.1)
        assert np.array_equal(w[b].cpu().numpy(), want["winding"])
        assert counts[b].tolist() == [want["num_vortices"], want["total_topological_charge"], want["abs_charge_count"]]
