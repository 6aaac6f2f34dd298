"""Neural closures (functions_nn.PeriodicCNN / Mixer2d, mirrors of pde_opt/numerics/functions/cnn.py:46-102 and
mixer_mlp.py:40-86) and the unfused-but-batched stepping path they take: closure evaluated on the whole batch,
stencils by pdeopt_rhs_given_mu_batched, spectral filter by pdeopt_sifs_filter_batched."""
import numpy as np
import pytest
import torch

from oracle import pde_oracle as O
from pde_opt_b200.functions import Mixer2d, PeriodicCNN, recognize

H, KAPPA = 0.01, 0.002


def test_periodic_cnn_is_translation_equivariant_and_batched():
    """cnn.py:49-51: 'Translation-equivariant on a torus'; [nx, ny] and [B, nx, ny] inputs give the same fields."""
    net = PeriodicCNN(1, (4, 8), kernel_size=3, key=3).double()
    x = torch.from_numpy(np.random.default_rng(0).normal(size=(2, 16, 16)))
    y = net(x)
    assert y.shape == x.shape and torch.allclose(net(x[1]), y[1], atol=1e-12)
    assert torch.allclose(net(torch.roll(x, (3, -5), (1, 2))), torch.roll(y, (3, -5), (1, 2)), atol=1e-12)
    assert recognize(net, "mu") is None  # a closure of the whole field: never mapped onto a pointwise family


def test_mixer2d_shapes_and_determinism():
    a = Mixer2d((1, 32, 32), 4, 8, 16, 16, 2, key=7)
    b = Mixer2d((1, 32, 32), 4, 8, 16, 16, 2, key=7)
    x = torch.from_numpy(np.random.default_rng(1).normal(size=(3, 32, 32)).astype(np.float32))
    assert a(x).shape == x.shape and torch.equal(a(x), b(x)) and a(x[0]).shape == (32, 32)


@pytest.mark.gpu
@pytest.mark.parametrize("n,kind,closure", [(128, "ch", "cnn"), (64, "ch", "mixer"), (128, "ac", "cnn"), (128, "ch", "cnn_mob")])
def test_unfused_step_with_neural_mu_matches_oracle(n, kind, closure):
    """One and sixteen semi-implicit steps with a network as mu (and, for 'cnn_mob', a non-enumerated callable as D):
    CUDA stencils + fused filter against the NumPy oracle evaluating the SAME network on the CPU."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import AllenCahn2DPeriodic, CahnHilliard2DPeriodic
    from pde_opt_b200.functions import ConstantMobility
    from pde_opt_b200.solvers import ODETerm, SemiImplicitFourierSpectral

    net = PeriodicCNN(1, (8, 8), kernel_size=3, key=11) if closure.startswith("cnn") else Mixer2d((1, n, n), 8, 8, 16, 16, 2, key=5)
    net_gpu = net.cuda()
    import copy

    net_cpu = copy.deepcopy(net).cpu().double()
    mob_gpu = (lambda c: 0.2 + 0.5 * torch.sigmoid(4.0 * c - 2.0)) if closure == "cnn_mob" else ConstantMobility(0.15)
    mob_np = (lambda c: (0.2 + 0.5 / (1.0 + np.exp(-(4.0 * c - 2.0)))).astype(c.dtype)) if closure == "cnn_mob" else (lambda c: 0.15 * np.ones_like(c))
    mu_np = lambda c: net_cpu(torch.from_numpy(np.asarray(c, np.float64))).detach().numpy().astype(np.float32)

    dom = Domain((n, n), ((-n * H / 2, n * H / 2),) * 2, "dimensionless")
    odom = O.Domain((n, n), ((-n * H / 2, n * H / 2),) * 2)
    if kind == "ch":
        eq, A, dt = CahnHilliard2DPeriodic(dom, KAPPA, net_gpu, mob_gpu), 0.5, 1e-6
        oeq = O.CahnHilliardPeriodic(odom, KAPPA, mu_np, mob_np, "fd", np.float32)
    else:
        eq, A, dt = AllenCahn2DPeriodic(dom, KAPPA, net_gpu, mob_gpu), 1.0, 5e-6
        oeq = O.AllenCahn2DPeriodic(odom, KAPPA, mu_np, mob_np, "fd", np.float32)
    assert not eq.fused
    solver = SemiImplicitFourierSpectral(A, eq.fourier_symbol, eq.fft, eq.ifft)
    B = 3
    y0 = np.stack([np.clip(0.5 + 0.1 * np.random.default_rng(i).normal(size=(n, n)), 0.05, 0.95) for i in range(B)]).astype(np.float32)
    yd = torch.from_numpy(y0).cuda()
    # right-hand side
    f = eq.rhs(yd).cpu().numpy()
    for b in range(B):
        want = oeq.rhs(y0[b], 0.0)
        assert np.linalg.norm(f[b] - want) <= 1e-4 * np.linalg.norm(want), (b, np.linalg.norm(f[b] - want) / np.linalg.norm(want))
    for K in (1, 16):
        times = np.arange(K + 1, dtype=np.float32) * np.float32(dt)
        got = solver.rollout(ODETerm(eq), times, yd).cpu().numpy()
        for b in range(B):
            y = y0[b]
            for k in range(K):
                y = O.sifs_step(oeq.rhs, y, times[k], times[k + 1], A, oeq.fourier_symbol)
            err = np.linalg.norm(got[b] - y) / np.linalg.norm(y)
            assert err <= 1e-5, (K, b, err)  # north star: 1e-5 after one step (held here after 16 as well)
            inc = np.linalg.norm((got[b] - y0[b]) - (y - y0[b])) / np.linalg.norm(y - y0[b])
            assert inc <= 2e-3, (K, b, inc)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ch", "ac"])
def test_gradients_through_neural_mu_match_float64_autograd(kind):
    """d loss / d (network parameters, y0) through 8 unfused steps: CUDA adjoint stencils + torch back-propagation through
    the network against torch.autograd on the float64 twin of the whole rollout (the same network in double)."""
    import copy

    from oracle import ch_torch_oracle as TO
    from pde_opt_b200 import Domain
    from pde_opt_b200.adjoint_nn import given_mu_rollout
    from pde_opt_b200.equations import AllenCahn2DPeriodic, CahnHilliard2DPeriodic
    from pde_opt_b200.functions import ConstantMobility
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    n, B, K = 64, 2, 8
    net = PeriodicCNN(1, (6, 6), kernel_size=3, key=2).cuda()
    net64 = copy.deepcopy(net).cpu().double()
    box = ((-n * H / 2, n * H / 2),) * 2
    dom = Domain((n, n), box, "dimensionless")
    if kind == "ch":
        eq, A, dt = CahnHilliard2DPeriodic(dom, KAPPA, net, ConstantMobility(0.15)), 0.5, 1e-6
    else:
        eq, A, dt = AllenCahn2DPeriodic(dom, KAPPA, net, ConstantMobility(0.15)), 1.0, 5e-6
    solver = SemiImplicitFourierSpectral(A, eq.fourier_symbol, eq.fft, eq.ifft)
    rng = np.random.default_rng(4)
    y0 = np.clip(0.5 + 0.1 * rng.normal(size=(B, n, n)), 0.05, 0.95).astype(np.float32)
    wgt = rng.normal(size=(B, n, n)).astype(np.float32)
    times = (np.arange(K + 1, dtype=np.float64) * dt).astype(np.float32)
    yg = torch.from_numpy(y0).cuda().requires_grad_(True)
    y1 = given_mu_rollout(eq, solver, yg, times)
    (y1 * torch.from_numpy(wgt).cuda()).sum().backward()

    y64 = torch.from_numpy(y0.astype(np.float64)).requires_grad_(True)
    dts = [float(d) for d in (times[1:] - times[:-1])]
    yr = TO.rollout(y64, dts, (n, n), box, KAPPA, A, lambda c: net64(c), lambda c: 0.15 * torch.ones_like(c), kind)
    (yr * torch.from_numpy(wgt.astype(np.float64))).sum().backward()
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    assert rel(y1.detach().cpu().numpy(), yr.detach().numpy()) <= 1e-5
    assert rel(yg.grad.cpu().numpy(), y64.grad.numpy()) <= 1e-4  # north star: gradients to relative 1e-4
    got = np.concatenate([p.grad.detach().cpu().numpy().ravel() for p in net.parameters()])
    want = np.concatenate([p.grad.detach().numpy().ravel() for p in net64.parameters()])
    assert rel(got, want) <= 1e-4


@pytest.mark.gpu
def test_train_mse_with_neural_mu_reduces_the_loss():
    """PDEModel.train(method="mse") with a PeriodicCNN as mu (optimization_neural_network.ipynb's set-up in miniature):
    data generated with the log potential from a smooth large-amplitude state (so that mu_h, not the gradient-energy term,
    drives the dynamics), a small pointwise network fitted to it; the loss goes down."""
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard2DPeriodic
    from pde_opt_b200.functions import ConstantMobility, LogRegular
    from pde_opt_b200.pde_model import PDEModel
    from pde_opt_b200.solvers import SemiImplicitFourierSpectral

    n = 64
    dom = Domain((n, n), ((-n * H / 2, n * H / 2),) * 2, "dimensionless")
    model = PDEModel(CahnHilliard2DPeriodic, dom, SemiImplicitFourierSpectral)
    X, Y = dom.mesh()
    L = n * H
    y0 = (0.5 + 0.25 * np.sin(2 * np.pi * X / L) * np.cos(4 * np.pi * Y / L) + 0.05 * np.cos(6 * np.pi * X / L)).astype(np.float32)[None]
    y0 = torch.from_numpy(y0).cuda()
    ts = np.array([0.0, 5e-4, 1e-3], np.float32)
    sol = model.solve({"kappa": KAPPA, "mu": LogRegular(3.0), "D": ConstantMobility(0.15)}, y0, ts, {"A": 0.5}, dt0=2e-5)
    assert float((sol[-1] - sol[0]).norm() / sol[0].norm()) > 5e-3  # the data carry a signal
    data = {"ys": [sol[t, 0].cpu().numpy() for t in range(3)], "ts": [float(t) for t in ts]}
    net = PeriodicCNN(1, (8,), kernel_size=1, key=1).cuda()  # pointwise network: the target is a pointwise function
    model.train(data, [[0, 1, 2]], {"mu": net}, {"kappa": KAPPA, "D": ConstantMobility(0.15)}, {"A": 0.5}, {}, 0.0, method="mse",
                max_steps=25, dt0=2e-5)
    h = model.last_loss_history
    assert h[-1] < 0.7 * h[0], h
