"""Advection-diffusion: fused forward kernel vs the NumPy oracle, hand-written adjoint vs
torch.autograd on the float64 oracle twin (BASELINE config 4; north-star tolerance 1e-4)."""
import numpy as np
import pytest
import torch

from oracle import ad_torch_oracle as TO
from oracle import pde_oracle as O

pytestmark = pytest.mark.gpu

N, H, DCOEF = 128, 0.02, 0.1
BOX = ((-N * H / 2, N * H / 2),) * 2


def _eq():
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AdvectionDiffusion2D
    from pde_opt_b200.functions import GaussianVelocity

    return AdvectionDiffusion2D(Domain((N, N), BOX, "dimensionless"), GaussianVelocity(0.1, 0.01), DCOEF)


def _y0(B, seed=0):
    return np.stack([0.5 + 0.01 * np.random.default_rng(seed + i).normal(size=(N, N)) for i in range(B)]).astype(np.float32)


def _ctrl(B, nseg, seed=1):
    rng = np.random.default_rng(seed)
    c = np.empty((B, nseg, 4), np.float32)
    c[..., 0] = rng.uniform(-0.5, 0.5, (B, nseg))
    c[..., 1] = rng.uniform(-0.5, 0.5, (B, nseg))
    c[..., 2] = rng.uniform(0.05, 0.2, (B, nseg))
    c[..., 3] = rng.uniform(0.01, 0.05, (B, nseg))
    return c


def _oracle_rollout(y0, ctrl, dts, hold):
    dom = O.Domain((N, N), BOX)
    out = np.empty_like(y0)
    for b in range(y0.shape[0]):
        y, t = y0[b], np.float32(0)
        for k, dt in enumerate(dts):
            s = min(k // hold, ctrl.shape[1] - 1)
            cx, cy, p0, p1 = (float(v) for v in ctrl[b, s])
            eq = O.AdvectionDiffusion2D(dom, O.gaussian_velocity((p0, p1), (cx, cy)), DCOEF, np.float32)
            y = O.sifs_step(eq.rhs, y, t, t + np.float32(dt), 1.0, eq.fourier_symbol)
            t = t + np.float32(dt)
        out[b] = y
    return out


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))


@pytest.mark.parametrize("B,K,nseg", [(2, 1, 1), (3, 16, 4), (1, 7, 7)])
def test_forward_matches_oracle(B, K, nseg):
    from pde_opt_b200.adjoint import ad_rollout

    eq = _eq()
    y0, ctrl = _y0(B), _ctrl(B, nseg)
    times = (np.arange(K + 1, dtype=np.float32) * np.float32(1e-4)).astype(np.float32)
    dts = times[1:] - times[:-1]
    hold = -(-K // nseg)
    got = ad_rollout(eq, torch.from_numpy(y0).cuda(), torch.from_numpy(ctrl).cuda(), times, hold=hold).cpu().numpy()
    want = _oracle_rollout(y0, ctrl, dts, hold)
    assert _rel(got, want) <= 1e-5  # north star: 1e-5 after one step
    # the increment itself (y1 - y0 is tiny next to y0, so test it separately)
    assert _rel(got - y0, want - y0) <= 2e-3


def test_forward_500_steps():
    from pde_opt_b200.adjoint import ad_rollout

    eq = _eq()
    B, K, nseg = 2, 500, 10
    y0, ctrl = _y0(B, 7), _ctrl(B, nseg, 3)
    times = (np.arange(K + 1, dtype=np.float32) * np.float32(1e-4)).astype(np.float32)
    got = ad_rollout(eq, torch.from_numpy(y0).cuda(), torch.from_numpy(ctrl).cuda(), times, hold=50).cpu().numpy()
    want = _oracle_rollout(y0, ctrl, times[1:] - times[:-1], 50)
    assert _rel(got, want) <= 1e-3  # north star: 1e-3 after 1000 steps
    assert _rel(got - y0, want - y0) <= 5e-3


def test_rhs_matches_oracle():
    eq = _eq()
    y0 = _y0(2, 11)
    f = eq.rhs(torch.from_numpy(y0).cuda()).cpu().numpy()
    dom = O.Domain((N, N), BOX)
    oeq = O.AdvectionDiffusion2D(dom, O.gaussian_velocity((0.1, 0.01), (0.0, 0.0)), DCOEF, np.float64)
    for b in range(2):
        want = oeq.rhs(y0[b].astype(np.float64))
        assert _rel(f[b], want) <= 5e-3  # (y1-y0)/dt in float32: cancellation-limited


@pytest.mark.parametrize("checkpoint_every", [None, 8])
def test_adjoint_matches_autograd(checkpoint_every):
    from pde_opt_b200.adjoint import ad_rollout

    eq = _eq()
    B, K, nseg = 3, 24, 3
    y0, ctrl = _y0(B, 21), _ctrl(B, nseg, 5)
    times = (np.arange(K + 1, dtype=np.float32) * np.float32(1e-4)).astype(np.float32)
    dts = (times[1:] - times[:-1]).astype(np.float64)
    wgt = np.random.default_rng(9).normal(size=(B, N, N)).astype(np.float32)

    yg = torch.from_numpy(y0).cuda().requires_grad_(True)
    cg = torch.from_numpy(ctrl).cuda().requires_grad_(True)
    yT = ad_rollout(eq, yg, cg, times, hold=8, checkpoint_every=checkpoint_every)
    loss = (yT * torch.from_numpy(wgt).cuda()).sum() + 0.5 * (yT**2).mean()
    loss.backward()

    yr = torch.from_numpy(y0.astype(np.float64)).requires_grad_(True)
    cr = torch.from_numpy(ctrl.astype(np.float64)).requires_grad_(True)
    yTr = TO.rollout(yr, cr, dts, (N, N), BOX, DCOEF, 1.0, hold=8)
    lr = (yTr * torch.from_numpy(wgt.astype(np.float64))).sum() + 0.5 * (yTr**2).mean()
    lr.backward()

    assert abs(loss.item() - lr.item()) <= 1e-5 * abs(lr.item())
    assert _rel(yg.grad.cpu().numpy(), yr.grad.numpy()) <= 1e-4
    gc, gr = cg.grad.cpu().numpy(), cr.grad.numpy()
    for j, name in enumerate(("cx", "cy", "p0", "p1")):
        assert _rel(gc[..., j], gr[..., j]) <= 1e-4, name


def test_adjoint_mean_square_loss_long():
    """SURVEY C4 loss (mean u^2 at T) over a longer rollout crossing the 512-step launch limit."""
    from pde_opt_b200.adjoint import ad_rollout

    eq = _eq()
    B, K, nseg = 2, 600, 6
    y0, ctrl = _y0(B, 31), _ctrl(B, nseg, 8)
    times = (np.arange(K + 1, dtype=np.float32) * np.float32(1e-4)).astype(np.float32)
    dts = (times[1:] - times[:-1]).astype(np.float64)
    yg = torch.from_numpy(y0).cuda().requires_grad_(True)
    cg = torch.from_numpy(ctrl).cuda().requires_grad_(True)
    loss = (ad_rollout(eq, yg, cg, times, hold=100) ** 2).mean()
    loss.backward()
    yr = torch.from_numpy(y0.astype(np.float64)).requires_grad_(True)
    cr = torch.from_numpy(ctrl.astype(np.float64)).requires_grad_(True)
    lr = (TO.rollout(yr, cr, dts, (N, N), BOX, DCOEF, 1.0, hold=100) ** 2).mean()
    lr.backward()
    assert _rel(yg.grad.cpu().numpy(), yr.grad.numpy()) <= 1e-4
    assert _rel(cg.grad.cpu().numpy(), cr.grad.numpy()) <= 1e-3  # float32 sums over 600 steps


def test_pde_model_mse_gradients_match_autograd_oracle():
    """PDEModel.mse (pde_model.py:274-323) on the advection-diffusion equation: loss and gradients
    w.r.t. the velocity parameters (the analogue of jax.grad(model.mse)) vs torch float64 autograd
    on the oracle twin, with save times that fall between step boundaries (interpolated)."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AdvectionDiffusion2D
    from pde_opt_b200.functions import GaussianVelocity
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    B = 3
    dom = Domain((N, N), BOX, "dimensionless")
    model = PDEModel(AdvectionDiffusion2D, dom, SemiImplicitFourierSpectral)
    y0 = _y0(B, 41)
    dt0 = 1e-4
    ts = np.asarray([0.0, 5e-4, 1.0e-3, 1.6e-3], dtype=np.float32)
    values = (y0[:, None] + 0.001 * np.random.default_rng(2).normal(size=(B, 3, N, N))).astype(np.float32)
    p0 = torch.tensor([0.10, 0.15, 0.08], device="cuda", requires_grad=True)
    p1 = torch.tensor([0.02, 0.03, 0.015], device="cuda", requires_grad=True)
    cx = torch.tensor([0.1, -0.2, 0.3], device="cuda", requires_grad=True)
    cy = torch.tensor([-0.1, 0.25, 0.0], device="cuda", requires_grad=True)
    params = {"velocity": GaussianVelocity(p0, p1, (cx, cy)), "D": DCOEF}
    loss = model.mse(params, (torch.from_numpy(y0).cuda(), torch.from_numpy(values).cuda()), {"A": 1.0}, ts, {}, 0.0, dt0=dt0)
    loss.backward()

    # oracle: float64 torch twin with the same float32 time grid and interpolation
    times = O.constant_step_schedule(ts[0], ts[-1], dt0, np.float32)
    yr = torch.from_numpy(y0.astype(np.float64))
    q = [torch.tensor(v.detach().cpu().numpy().astype(np.float64), requires_grad=True) for v in (cx, cy, p0, p1)]
    ctrl = torch.stack(q, -1)[:, None, :]
    hold = len(times) - 1
    preds, y, i_cur = [], yr, 0

    def seg(y, a, b):
        d = (times[a + 1 : b + 1] - times[a:b]).astype(np.float64)
        return TO.rollout(y, ctrl, d, (N, N), BOX, DCOEF, 1.0, hold=hold)

    for s_ in ts:
        j = int(np.searchsorted(times, s_, side="left"))
        if j == 0 or times[j] == s_:
            if j > i_cur:
                y, i_cur = seg(y, i_cur, j), j
            preds.append(y)
            continue
        if j - 1 > i_cur:
            y, i_cur = seg(y, i_cur, j - 1), j - 1
        yb = seg(y, j - 1, j)
        w = float(np.float32((s_ - times[j - 1]) / (times[j] - times[j - 1])))
        preds.append(y + (yb - y) * w)
        y, i_cur = yb, j
    pred = torch.stack(preds, 0)
    lr = ((torch.from_numpy(values.astype(np.float64)) - pred[1:].transpose(0, 1)) ** 2).mean()
    lr.backward()
    assert abs(loss.item() - lr.item()) <= 1e-4 * abs(lr.item())
    for got, want, name in zip((cx, cy, p0, p1), q, ("cx", "cy", "p0", "p1")):
        assert _rel(got.grad.cpu().numpy(), want.grad.numpy()) <= 2e-4, name


def test_reference_fixture_notebooks_reference_npy_on_gpu():
    """The reference's own advection-diffusion result (notebooks/reference.npy, 64x64, t = 5; committed
    as tests/golden/ref_advection_diffusion_64.npy) reproduced by the GPU path in float32:
    2500 semi-implicit steps of dt = 2e-3 from the (decayed-noise) mean state."""
    import os

    from pde_opt_b200 import Domain
    from pde_opt_b200.adjoint import ad_rollout
    from pde_opt_b200.equations import AdvectionDiffusion2D
    from pde_opt_b200.functions import GaussianVelocity

    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_advection_diffusion_64.npy"))
    n, L = 64, 0.02 * 64
    dom = Domain((n, n), ((-L / 2, L / 2), (-L / 2, L / 2)), "dimensionless")
    eq = AdvectionDiffusion2D(dom, GaussianVelocity(0.1, 0.01, (0.4, 0.4)), 0.1)
    y0 = torch.full((1, n, n), float(ref.mean()), dtype=torch.float32, device="cuda")
    times = (np.arange(2501, dtype=np.float64) * 2e-3).astype(np.float32)
    ctrl = eq.control_block(1, "cuda")
    got = ad_rollout(eq, y0, ctrl, times)[0].cpu().numpy()
    assert _rel(got, ref) < 2e-4


@pytest.mark.parametrize("shape", [(64, 64), (32, 32), (32, 128), (16, 16)])
def test_generic_sizes_match_oracle(shape):
    from pde_opt_b200 import Domain
    from pde_opt_b200.adjoint import ad_rollout
    from pde_opt_b200.equations import AdvectionDiffusion2D
    from pde_opt_b200.functions import GaussianVelocity

    nx, ny = shape
    box = ((-nx * H / 2, nx * H / 2), (-ny * H / 2, ny * H / 2))
    eq = AdvectionDiffusion2D(Domain(shape, box, "dimensionless"), GaussianVelocity(0.1, 0.01), DCOEF)
    B, K, nseg = 3, 6, 2
    rng = np.random.default_rng(4)
    y0 = (0.5 + 0.01 * rng.normal(size=(B,) + shape)).astype(np.float32)
    ctrl = _ctrl(B, nseg, 6)
    ctrl[..., :2] *= 0.3
    times = (np.arange(K + 1, dtype=np.float32) * np.float32(1e-4)).astype(np.float32)
    got = ad_rollout(eq, torch.from_numpy(y0).cuda(), torch.from_numpy(ctrl).cuda(), times, hold=3).cpu().numpy()
    dom = O.Domain(shape, box)
    for b in range(B):
        y, t = y0[b], np.float32(0)
        for k in range(K):
            cx, cy, p0, p1 = (float(v) for v in ctrl[b, min(k // 3, nseg - 1)])
            oeq = O.AdvectionDiffusion2D(dom, O.gaussian_velocity((p0, p1), (cx, cy)), DCOEF, np.float32)
            y = O.sifs_step(oeq.rhs, y, t, t + np.float32(1e-4), 1.0, oeq.fourier_symbol)
            t = t + np.float32(1e-4)
        assert _rel(got[b], y) <= 1e-5
        assert _rel(got[b] - y0[b], y - y0[b]) <= 2e-3


def test_pde_model_train_recovers_velocity_parameters():
    """PDEModel.train(method="mse") (pde_model.py:325-460): synthetic trajectories generated with known
    velocity parameters; starting from perturbed values the quasi-Newton loop over the adjoint
    gradients drives the loss down by orders of magnitude and recovers the parameters."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AdvectionDiffusion2D
    from pde_opt_b200.functions import GaussianVelocity
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    dom = Domain((N, N), BOX, "dimensionless")
    model = PDEModel(AdvectionDiffusion2D, dom, SemiImplicitFourierSpectral)
    xs, ys_ = np.meshgrid(*[np.linspace(b[0] + H / 2, b[1] - H / 2, N) for b in BOX], indexing="ij")
    u0 = (0.5 + 0.2 * np.exp(-((xs - 0.2) ** 2 + (ys_ + 0.1) ** 2) / 0.05)).astype(np.float32)
    true = dict(p0=0.5, p1=0.05, cx=0.15, cy=-0.2)
    ts = [0.0, 0.01, 0.02, 0.03]
    truth = model.solve({"velocity": GaussianVelocity(true["p0"], true["p1"], (true["cx"], true["cy"])), "D": DCOEF},
                        torch.from_numpy(u0).cuda(), ts, {"A": 1.0}, dt0=2e-4)
    data = {"ys": [truth[i].detach().cpu().numpy() for i in range(4)], "ts": ts}
    p0 = torch.tensor(0.3, device="cuda")
    cx = torch.tensor(0.05, device="cuda")
    cy = torch.tensor(-0.1, device="cuda")
    opt = {"velocity": GaussianVelocity(p0, true["p1"], (cx, cy))}
    out = model.train(data, [[0, 1, 2, 3]], opt, {"D": DCOEF}, {"A": 1.0}, {}, 0.0, method="mse", max_steps=40, dt0=2e-4)
    hist = model.last_loss_history
    assert hist[-1] < 1e-4 * hist[0]
    assert abs(p0.item() - true["p0"]) < 0.02 * true["p0"]
    assert abs(cx.item() - true["cx"]) < 5e-3 and abs(cy.item() - true["cy"]) < 5e-3
    assert out["D"] == DCOEF and out["velocity"] is opt["velocity"]
    with pytest.raises(NotImplementedError):
        model.train(data, [[0, 1]], opt, {"D": DCOEF}, {"A": 1.0}, {}, 0.0, method="least_squares")
