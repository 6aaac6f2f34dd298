import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import ch_torch_oracle as TO
from pde_opt_b200 import Domain
from pde_opt_b200.adjoint_ch import phasefield_rollout
from pde_opt_b200.equations import CahnHilliard2DPeriodic
from pde_opt_b200.functions import ChemicalPotentialLegendrePolynomials, DiffusionLegendrePolynomials
from pde_opt_b200.solvers import SemiImplicitFourierSpectral
n, H, KAPPA, K = 64, 0.01, 0.002, 2000
box = ((0.0, n * H), (0.0, n * H))
mu_t = torch.tensor([0.1, 2.5, -0.3, 0.8], device="cuda", requires_grad=True)
mob_t = torch.tensor([-1.0, 0.3, -0.2], device="cuda", requires_grad=True)
eq = CahnHilliard2DPeriodic(Domain((n, n), box, "d"), KAPPA, ChemicalPotentialLegendrePolynomials(mu_t, "log"), DiffusionLegendrePolynomials(mob_t))
solver = SemiImplicitFourierSpectral(0.5, eq.fourier_symbol, eq.fft, eq.ifft)
rng = np.random.default_rng(1)
y0 = np.clip(0.5 + 0.05 * rng.normal(size=(1, n, n)), 0.1, 0.9).astype(np.float32)
times = (np.arange(K + 1, dtype=np.float64) * 1e-6).astype(np.float32)
yg = torch.from_numpy(y0).cuda().requires_grad_(True)
yT = phasefield_rollout(eq, solver, yg, times)
loss = (yT**2).mean(); loss.backward()
yr = torch.from_numpy(y0.astype(np.float64)).requires_grad_(True)
pm = torch.tensor(mu_t.detach().cpu().numpy().astype(np.float64), requires_grad=True)
pd = torch.tensor(mob_t.detach().cpu().numpy().astype(np.float64), requires_grad=True)
dts = (times[1:] - times[:-1]).astype(np.float64)
yTr = TO.rollout(yr, dts, (n, n), box, KAPPA, 0.5, lambda c: TO.mu_legendre(pm, c, True), lambda c: TO.D_legendre(pd, c), "ch")
lr = (yTr**2).mean(); lr.backward()
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
print("state", rel(yT.detach().cpu().numpy(), yTr.detach().numpy()), "loss", abs(loss.item() - lr.item()) / abs(lr.item()))
print("gy0", rel(yg.grad.cpu().numpy(), yr.grad.numpy()), "gmu", rel(mu_t.grad.cpu().numpy(), pm.grad.numpy()), "gD", rel(mob_t.grad.cpu().numpy(), pd.grad.numpy()))
