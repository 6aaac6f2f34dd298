"""Forward-mode tangents (pdeopt_phasefield_tangent_steps), the checkpointing forward (pdeopt_sifs_rollout_fwd) and
PDEModel.train(method="least_squares") — the reference's default training method (pde_model.py:334,404-428).
Oracle: torch float64 twin of the reference arithmetic (oracle/ch_torch_oracle.py), Jacobian-vector products by
torch.autograd.functional.jvp (stand-in for diffrax's ForwardMode)."""
import numpy as np
import pytest
import torch

from oracle import ch_torch_oracle as TO
from pde_opt_b200 import Domain, _lib
from pde_opt_b200.equations import AllenCahn2DPeriodic, CahnHilliard2DPeriodic
from pde_opt_b200.functions import ChemicalPotentialLegendrePolynomials, ConstantMobility, DiffusionLegendrePolynomials
from pde_opt_b200.pde_model import PDEModel
from pde_opt_b200.solvers import SemiImplicitFourierSpectral

pytestmark = pytest.mark.gpu

KAPPA, H = 0.002, 0.01
MU_TRUE = [0.1, -2.2, 0.3, 0.25]
D_TRUE = [-0.3, 0.2, -0.1]


def _dom(n):
    return Domain((n, n), ((-n * H / 2, n * H / 2),) * 2, "dimensionless")


def _ic(n, B, seed=0, amp=0.1):
    rng = np.random.default_rng(seed)
    return np.clip(0.5 + amp * rng.normal(size=(B, n, n)), 0.05, 0.95).astype(np.float32)


@pytest.mark.parametrize("n", [128, 64])
def test_rollout_fwd_checkpoints(n):
    """The states the rollout keeps are the states of K single-step launches: bit for bit on the fused 128 x 128 kernel;
    on the small-grid kernel a 3-step launch and three 1-step launches differ by float32 rounding."""
    eq = CahnHilliard2DPeriodic(_dom(n), KAPPA, ChemicalPotentialLegendrePolynomials(MU_TRUE, "log"), DiffusionLegendrePolynomials(D_TRUE))
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    y0 = torch.from_numpy(_ic(n, 3)).cuda()
    dts = np.full(7, 1e-6, np.float32)
    plan, sym = eq.plan(), solver.symbol_on(y0.device)
    y1, traj = plan.rollout_fwd(y0, dts, sym, save_every=1)
    y = y0.clone()
    for k in range(7):
        assert torch.equal(traj[k], y), k
        y = plan.step(y, dts[k:k + 1], sym)
    assert torch.equal(y1, y)
    y1b, ck = plan.rollout_fwd(y0, dts, sym, save_every=3)
    same = torch.equal if n == 128 else (lambda a, b: torch.allclose(a, b, rtol=1e-5, atol=1e-7))
    assert ck.shape[0] == 3 and torch.equal(ck[0], traj[0]) and same(ck[1], traj[3]) and same(ck[2], traj[6])
    assert same(y1b, y1)


@pytest.mark.parametrize("kind,n,K", [("ch", 128, 40), ("ch", 64, 25), ("ac", 128, 40)])
def test_tangent_vs_oracle_jvp(kind, n, K):
    """d y_K / d theta (every Legendre coefficient of mu and of the mobility) within 1e-4 of the float64 JVP."""
    dom = _dom(n)
    mu = ChemicalPotentialLegendrePolynomials(torch.tensor(MU_TRUE), "log")
    mob = DiffusionLegendrePolynomials(torch.tensor(D_TRUE))
    if kind == "ch":
        eq, A, dt = CahnHilliard2DPeriodic(dom, KAPPA, mu, mob), 0.5, 1e-6
    else:
        eq, A, dt = AllenCahn2DPeriodic(dom, KAPPA, mu, mob), 1.0, 1e-4
    solver = SemiImplicitFourierSpectral(A, eq.fourier_symbol, eq.fft, eq.ifft)
    y0 = _ic(n, 2, seed=3)
    dts = np.full(K, dt, np.float32)
    plan, sym = eq.plan(), solver.symbol_on("cuda")
    yd = torch.from_numpy(y0).cuda()
    _, traj = plan.rollout_fwd(yd, dts, sym, save_every=1)
    ndir = len(MU_TRUE) + len(D_TRUE)
    dmu = torch.zeros((ndir, _lib.MAX_COEF), device="cuda")
    dmob = torch.zeros_like(dmu)
    for i in range(len(MU_TRUE)):
        dmu[i, i] = 1.0
    for i in range(len(D_TRUE)):
        dmob[len(MU_TRUE) + i, i] = 1.0
    v = torch.zeros((ndir, 2, n, n), device="cuda")
    plan.tangent_steps(traj, v, dts, dmu, dmob, sym)
    got = v.cpu().double().numpy()

    box = ((-n * H / 2, n * H / 2),) * 2
    y64 = torch.from_numpy(y0.astype(np.float64))

    def f(pm, pd):
        return TO.rollout(y64, [float(d) for d in dts], (n, n), box, KAPPA, A, lambda c: TO.mu_legendre(pm, c, True),
                          lambda c: TO.D_legendre(pd, c), kind)

    pm, pd = torch.tensor(MU_TRUE, dtype=torch.float64), torch.tensor(D_TRUE, dtype=torch.float64)
    for d in range(ndir):
        tm, td = torch.zeros_like(pm), torch.zeros_like(pd)
        if d < len(MU_TRUE):
            tm[d] = 1.0
        else:
            td[d - len(MU_TRUE)] = 1.0
        _, want = torch.autograd.functional.jvp(f, (pm, pd), (tm, td))
        want = want.numpy()
        # (the constant coefficient of mu does not enter the Cahn-Hilliard dynamics: both tangents are exactly zero)
        err = np.linalg.norm(got[d] - want) / max(np.linalg.norm(want), 1e-300)
        assert err <= 1e-4, (kind, n, d, err)  # tolerance: north star, gradients to relative 1e-4


def test_state_tangent_matches_finite_difference():
    """A state-only tangent (dtheta = 0, v0 != 0) equals the directional finite difference of the GPU rollout."""
    n, K = 128, 8
    eq = CahnHilliard2DPeriodic(_dom(n), KAPPA, ChemicalPotentialLegendrePolynomials(MU_TRUE, "log"), ConstantMobility(0.15))
    solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
    plan, sym = eq.plan(), solver.symbol_on("cuda")
    dts = np.full(K, 1e-6, np.float32)
    y0 = torch.from_numpy(_ic(n, 1, seed=5)).cuda()
    rng = np.random.default_rng(9)
    w = torch.from_numpy(rng.normal(size=(1, n, n)).astype(np.float32)).cuda()
    y1, traj = plan.rollout_fwd(y0, dts, sym, save_every=1)
    v = w.clone().unsqueeze(0).contiguous()
    z = torch.zeros((1, _lib.MAX_COEF), device="cuda")
    plan.tangent_steps(traj, v, dts, z, z, sym)
    eps = 1e-2
    yp = plan.step((y0 + eps * w).contiguous(), dts, sym)
    ym = plan.step((y0 - eps * w).contiguous(), dts, sym)
    fd = (yp - ym) / (2 * eps)
    err = float((v[0] - fd).norm() / fd.norm())
    assert err <= 2e-3, err


def test_levenberg_marquardt_recovers_legendre_coefficients():
    """Synthetic Cahn-Hilliard trajectories generated with known Legendre coefficients of mu; train() with the
    reference's default method starts from perturbed coefficients and recovers them."""
    n = 64
    dom = _dom(n)
    model = PDEModel(CahnHilliard2DPeriodic, dom, SemiImplicitFourierSpectral)
    true = {"kappa": KAPPA, "mu": ChemicalPotentialLegendrePolynomials(MU_TRUE, "log"), "D": ConstantMobility(0.15)}
    ts = np.array([0.0, 2e-5, 4e-5, 6e-5], np.float32)
    y0 = torch.from_numpy(_ic(n, 2, seed=11, amp=0.15)).cuda()
    sol = model.solve(true, y0, ts, {"A": 0.5}, dt0=1e-6)  # [T, B, n, n]
    # two trajectories, each a list of snapshots (the reference's data layout, pde_model.py:381-396)
    data = {"ys": [sol[t, b].cpu().numpy() for b in range(2) for t in range(4)], "ts": [float(t) for t in ts] * 2}
    inds = [[0, 1, 2, 3], [4, 5, 6, 7]]
    start = torch.tensor([0.0, -2.0, 0.0, 0.0])
    opt = {"mu": ChemicalPotentialLegendrePolynomials(start.clone().cuda(), "log")}
    out = model.train(data, inds, opt, {"kappa": KAPPA, "D": ConstantMobility(0.15)}, {"A": 0.5}, {}, 0.0, max_steps=30, dt0=1e-6)
    got = out["mu"].coef.detach().cpu().numpy()
    hist = model.last_loss_history
    assert hist[-1] <= 1e-6 * hist[0], hist
    # mu enters the dynamics only through its gradient: the constant coefficient is not identifiable
    assert np.allclose(got[1:], MU_TRUE[1:], atol=2e-3), got
    # regularisation row: the reference appends reg = lambda sum w theta^2 to the residual pytree (pde_model.py:226-272),
    # so Levenberg-Marquardt sees 1/2 reg^2 — a quartic pull on the (otherwise unidentifiable) constant coefficient
    opt = {"mu": ChemicalPotentialLegendrePolynomials(torch.tensor([0.5, -2.0, 0.0, 0.0]).cuda(), "log")}
    w = {"mu": torch.tensor([1.0, 0.0, 0.0, 0.0])}
    out = model.train(data, inds, opt, {"kappa": KAPPA, "D": ConstantMobility(0.15)}, {"A": 0.5}, w, 1.0, max_steps=30, dt0=1e-6)
    got = out["mu"].coef.detach().cpu().numpy()
    assert abs(got[0]) < 0.05 and np.allclose(got[1:], MU_TRUE[1:], atol=1e-2), got


def test_least_squares_rejects_unsupported_leaves():
    n = 64
    model = PDEModel(CahnHilliard2DPeriodic, _dom(n), SemiImplicitFourierSpectral)
    y = _ic(n, 1)[0]
    data = {"ys": [y, y], "ts": [0.0, 1e-6]}
    with pytest.raises(NotImplementedError):
        model.train(data, [[0, 1]], {"kappa": torch.tensor(KAPPA)}, {"mu": ChemicalPotentialLegendrePolynomials(MU_TRUE, "log"),
                    "D": ConstantMobility(0.15)}, {"A": 0.5}, {}, 0.0, max_steps=2)
