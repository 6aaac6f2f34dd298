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
