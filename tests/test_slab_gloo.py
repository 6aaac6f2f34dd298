"""world_size-2 gloo test (CPU) of the slab-decomposed 3-D Cahn-Hilliard step: halo exchange, packed
line geometries and the two all-to-alls, with the four device operations emulated in NumPy (the
oracle's stencils and numpy.fft) so that the distributed choreography is what is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pde_oracle as O

POINTS, H, KAPPA, A, DT = (16, 16, 8), 0.01, 0.002, 0.5, 1e-6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _offsets(g, n):
    line = np.arange(g.n_lines)[:, None]
    idx = np.arange(n)[None, :]
    return (line // g.n_inner) * g.outer + (line % g.n_inner) * g.inner + (idx // g.chunk) * g.hi + (idx % g.chunk) * g.lo


class EmuBackend:
    """NumPy emulation of pdeopt_ch3d_rhs / pdeopt_fft_lines* (semantics of include/pdeopt_b200.h)."""

    def __init__(self, h):
        from pde_opt_b200.linefft import pos_to_freq

        self.h, self.p2f = h, pos_to_freq

    def rhs(self, u, lo, hi):
        ext = np.concatenate([lo.numpy(), u.numpy(), hi.numpy()], 0)
        dom = O.Domain(ext.shape, tuple((0.0, n * self.h) for n in ext.shape))
        eq = O.CahnHilliardPeriodic(dom, KAPPA, lambda c: O.mu_log(c, 3.0), lambda c: 0.15 * np.ones_like(c), "fd", np.float32)
        return torch.from_numpy(np.ascontiguousarray(eq.rhs_fd(ext)[2:-2]))

    def fft_r2c(self, f, dst, n, n_lines):
        lines = f.numpy().reshape(n_lines, n)
        dst.numpy().ravel()[: n_lines * (n // 2 + 1)] = np.fft.rfft(lines, axis=1).astype(np.complex64).ravel()

    def fft_c2r_update(self, spec, n, n_lines, y0, y1, dt):
        half = spec.numpy().ravel()[: n_lines * (n // 2 + 1)].reshape(n_lines, n // 2 + 1)
        g = (np.fft.irfft(half, n=n, axis=1) * n).astype(np.float32)
        y1.numpy().reshape(n_lines, n)[:] = y0.numpy().reshape(n_lines, n) + np.float32(dt) * g

    def fft_lines(self, src, dst, n, gin, gout, inverse, in_real, scale):
        a = src.numpy().ravel()
        lines = a[_offsets(gin, n)].astype(np.complex64)
        if inverse:
            nat = np.empty_like(lines)
            nat[:, self.p2f(n)] = lines
            res = np.fft.ifft(nat, axis=1) * n
        else:
            res = np.fft.fft(lines, axis=1)[:, self.p2f(n)]
        dst.numpy().ravel()[_offsets(gout, n)] = (res * scale).astype(np.complex64)

    def fft_lines_imex(self, buf, n, g, sym, gsym, dt, scale):
        a = buf.numpy().ravel()
        off = _offsets(g, n)
        spec = np.fft.fft(a[off].astype(np.complex64), axis=1)[:, self.p2f(n)]
        spec = spec * (np.float32(scale) / (np.float32(1) + np.float32(dt) * sym.numpy().ravel()[_offsets(gsym, n)]))
        nat = np.empty_like(spec)
        nat[:, self.p2f(n)] = spec
        a[off] = (np.fft.ifft(nat, axis=1) * n).astype(np.complex64)

    def fft_lines_inv_update(self, spec, n, gin, y0, y1, gout, dt):
        lines = spec.numpy().ravel()[_offsets(gin, n)]
        nat = np.empty_like(lines)
        nat[:, self.p2f(n)] = lines
        g = (np.fft.ifft(nat, axis=1) * n).real.astype(np.float32)
        off = _offsets(gout, n)
        y1.numpy().ravel()[off] = y0.numpy().ravel()[off] + np.float32(dt) * g


def _global_problem():
    from pde_opt_b200 import Domain
    from pde_opt_b200.equations import CahnHilliard3DPeriodic
    from pde_opt_b200.functions import ConstantMobility, LogRegular

    box = tuple((0.0, n * H) for n in POINTS)
    eq = CahnHilliard3DPeriodic(Domain(POINTS, box, "dimensionless"), KAPPA, LogRegular(3.0), ConstantMobility(0.15))
    u0 = np.clip(0.5 + 0.01 * np.random.default_rng(0).normal(size=POINTS), 0.01, 0.99).astype(np.float32)
    return eq, box, u0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pde_opt_b200.parallel import SlabCahnHilliard3D

    eq, box, u0 = _global_problem()
    slab = SlabCahnHilliard3D(eq, A, backend=EmuBackend(H), device="cpu")
    nxl = POINTS[0] // world
    u = torch.from_numpy(u0[rank * nxl : (rank + 1) * nxl].copy())
    for _ in range(2):
        u = slab.step(u, DT)
    q.put((rank, u.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _oracle_two_steps():
    eq, box, u0 = _global_problem()
    oeq = O.CahnHilliardPeriodic(O.Domain(POINTS, box), KAPPA, lambda c: O.mu_log(c, 3.0), lambda c: 0.15 * np.ones_like(c), "fd", np.float32)
    y = u0
    for k in range(2):
        y = O.sifs_step(oeq.rhs, y, np.float32(k * DT), np.float32((k + 1) * DT), A, oeq.fourier_symbol)
    return u0, y


def test_slab_step_world2_matches_global_oracle():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.concatenate([res[r] for r in range(world)], 0)
    u0, want = _oracle_two_steps()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-6
    assert np.linalg.norm((got - u0) - (want - u0)) / np.linalg.norm(want - u0) <= 1e-3


def test_slab_step_single_rank_emulation():
    """world_size 1 (no process group): same code path with copies instead of collectives."""
    from pde_opt_b200.parallel import SlabCahnHilliard3D

    eq, box, u0 = _global_problem()
    slab = SlabCahnHilliard3D(eq, A, backend=EmuBackend(H), device="cpu")
    u = torch.from_numpy(u0.copy())
    for _ in range(2):
        u = slab.step(u, DT)
    _, want = _oracle_two_steps()
    assert np.linalg.norm(u.numpy() - want) / np.linalg.norm(want) <= 1e-6
