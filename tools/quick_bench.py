"""Quick device-only timing of the fused CH kernel (not the contract bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pde_opt_b200.fused import SifsPlan

N, H, KAPPA = 128, 0.01, 0.002
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
fam = sys.argv[3] if len(sys.argv) > 3 else "log_degenerate"
mu, mob = {"log_degenerate": (("log", (3.0,)), ("degenerate", ())), "dw_const": (("double_well", ()), ("const", (1.0,)))}[fam]
plan = SifsPlan("ch2d", N, N, (-N * H / 2,) * 2, (H, H), KAPPA, mu, mob)
k = np.fft.fftfreq(N, H)[: N // 2 + 1]
k2 = (2 * np.pi) ** 2 * (k[:, None] ** 2 + k[None, :] ** 2)
quad = (KAPPA * k2 ** 2).astype(np.float32)
tab = torch.from_numpy(np.ascontiguousarray(0.5 * quad)).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
y = (0.5 + 0.01 * torch.randn((B, N, N), device="cuda", generator=g)).clamp(0, 1).contiguous()
out = torch.empty_like(y)
dts = [1e-6] * K
for _ in range(3):
    plan.step(y, dts, tab, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    plan.step(y, dts, tab, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"B={B} K={K} {fam}: {ms:.3f} ms/launch  {B * K / ms * 1e3 / 1e6:.2f} M env-steps/s  finite={bool(torch.isfinite(out).all())}")
