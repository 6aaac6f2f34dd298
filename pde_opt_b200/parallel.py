"""Multi-GPU plumbing: environment batches are sharded over ranks (one process per GPU) with
no communication inside a step; the only collective on the env path is the all-gather of the
per-environment rewards / losses.  Works with the `nccl` backend on GPUs and with
`gloo` on CPU tensors (used by the world_size-2 tests)."""
import torch
import torch.distributed as dist


def shard_bounds(num_envs, rank, world):
    """Contiguous, balanced [lo, hi) slice of the env axis owned by `rank` (first `num_envs %
    world` ranks hold one extra environment)."""
    base, extra = divmod(int(num_envs), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_env_values(local, num_envs=None, group=None):
    """All-gather a per-environment tensor [b_local, ...] into [num_envs, ...] on every rank.
    Shards may be ragged (see shard_bounds): they are padded to the largest shard for the
    collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if num_envs is None:
        n = torch.tensor([local.shape[0]], device=local.device)
        dist.all_reduce(n, group=group)
        num_envs = int(n.item())
    sizes = [shard_bounds(num_envs, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((bmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * bmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * bmax : r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    del rank
    return torch.cat(parts, dim=0)


# --------------------------------------------------------------------------------------------------
# Slab-decomposed 3-D Cahn-Hilliard (BASELINE config 5)
# --------------------------------------------------------------------------------------------------
class _DeviceBackend:
    """The device operations of the slab step, through the C ABI (include/pdeopt_b200.h)."""

    def __init__(self, plan):
        self.plan = plan

    def rhs(self, u, halo_lo, halo_hi):
        return self.plan.rhs(u[None], halo_lo, halo_hi)[0]

    @staticmethod
    def _stream(t):
        import ctypes

        return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    def fft_r2c(self, f, dst, n, n_lines):
        import ctypes

        from . import _lib

        _lib.check(_lib.load().pdeopt_fft_lines_r2c(ctypes.c_void_p(f.data_ptr()), ctypes.c_void_p(dst.data_ptr()), n, n_lines,
                                                   self._stream(f)))

    def fft_c2r_update(self, spec, n, n_lines, y0, y1, dt):
        import ctypes

        from . import _lib

        _lib.check(_lib.load().pdeopt_fft_lines_c2r_update(ctypes.c_void_p(spec.data_ptr()), n, n_lines,
                                                          ctypes.c_void_p(y0.data_ptr()), ctypes.c_void_p(y1.data_ptr()),
                                                          float(dt), self._stream(spec)))

    def fft_lines(self, src, dst, n, gin, gout, inverse, in_real, scale):
        import ctypes

        from . import _lib

        _lib.check(_lib.load().pdeopt_fft_lines(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), n,
                                               ctypes.byref(gin), ctypes.byref(gout), int(inverse), int(in_real), float(scale),
                                               self._stream(src)))

    def fft_lines_to_peers(self, src, n, gin, peer_ptrs, gout, src_off, sym=None, gsym=None, dt=0.0, scale=1.0):
        import ctypes

        from . import _lib

        arr = (ctypes.c_void_p * len(peer_ptrs))(*[ctypes.c_void_p(int(p)) for p in peer_ptrs])
        _lib.check(_lib.load().pdeopt_fft_lines_to_peers(
            ctypes.c_void_p(src.data_ptr()), n, ctypes.byref(gin), arr, len(peer_ptrs), ctypes.byref(gout), int(src_off),
            ctypes.c_void_p(sym.data_ptr()) if sym is not None else None, ctypes.byref(gsym) if gsym is not None else None,
            float(dt), float(scale), self._stream(src)))

    def push_blocks(self, src, peer_ptrs, block_bytes, dst_off_bytes, first):
        import ctypes

        from . import _lib

        arr = (ctypes.c_void_p * len(peer_ptrs))(*[ctypes.c_void_p(int(p)) for p in peer_ptrs])
        _lib.check(_lib.load().pdeopt_push_blocks_to_peers(ctypes.c_void_p(src.data_ptr()), arr, len(peer_ptrs), int(block_bytes),
                                                           int(dst_off_bytes), int(first), self._stream(src)))

    def push_rows(self, src, peer_ptrs, src_block_bytes, dst_off_bytes, n_rows, run_bytes, row_stride_bytes, first):
        import ctypes

        from . import _lib

        key = tuple(int(p) for p in peer_ptrs)
        cache = self.__dict__.setdefault("_arr_cache", {})
        arr = cache.get(key)
        if arr is None:
            arr = cache[key] = (ctypes.c_void_p * len(peer_ptrs))(*[ctypes.c_void_p(p) for p in key])
        _lib.check(_lib.load().pdeopt_push_rows_to_peers(ctypes.c_void_p(src.data_ptr()), arr, len(peer_ptrs), int(src_block_bytes),
                                                         int(dst_off_bytes), int(n_rows), int(run_bytes), int(row_stride_bytes), int(first),
                                                         self._stream(src)))

    def fft_lines_imex(self, buf, n, g, sym, gsym, dt, scale):
        import ctypes

        from . import _lib

        _lib.check(_lib.load().pdeopt_fft_lines_imex(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(buf.data_ptr()), n,
                                                    ctypes.byref(g), ctypes.c_void_p(sym.data_ptr()), ctypes.byref(gsym),
                                                    float(dt), float(scale), self._stream(buf)))

    def fft_lines_inv_update(self, spec, n, gin, y0, y1, gout, dt):
        import ctypes

        from . import _lib

        _lib.check(_lib.load().pdeopt_fft_lines_inv_update(ctypes.c_void_p(spec.data_ptr()), n, ctypes.byref(gin),
                                                          ctypes.c_void_p(y0.data_ptr()), ctypes.c_void_p(y1.data_ptr()),
                                                          ctypes.byref(gout), float(dt), self._stream(spec)))


class SlabCahnHilliard3D:
    """One semi-implicit step (solvers.py:56-70) of CahnHilliard3DPeriodic.rhs_fd
    (cahn_hilliard.py:177-200) on a single large domain sharded over ranks by x-slabs.

    Rank r holds planes [r*nxl, (r+1)*nxl) of the global [Nx, Ny, Nz] field.  Per step:
      1. ring exchange of the two boundary planes on either side (the FD RHS reaches x +- 2),
      2. rhs_fd on the slab,
      3. real-to-half-spectrum z transform (Hz = Nz/2+1 complex per line), then the y line FFTs; the
         y pass writes straight into the packed send buffer [dest][nxl][Ny/P][Hz] (chunked line
         geometry: no separate pack kernel),
      4. all-to-all #1 -> [Nx][Ny/P][Hz]: x lines complete on every rank,
      5. x pass: forward FFT, multiply by 1/(N (1 + A dt sigma)), inverse FFT in one kernel,
      6. all-to-all #2 back (no pack needed: x-ranges are contiguous),
      7. inverse y pass reading the packed layout, half-spectrum-to-real z pass fused with y1 = y0 + dt g.
    With transport="peer" (GPUs of one NVSwitch domain) the two all-to-alls are not separate
    collectives: the y pass and the x pass store their last stage straight into the peers' buffers
    (torch symmetric memory, P2P over NVLink; pdeopt_fft_lines_to_peers), so the transpose rides on the
    transform kernels and only a cross-rank barrier separates producer and consumer.
    The spectrum stays in the line engine's position order throughout; the symbol table is permuted
    once at construction.  Collectives: torch.distributed (NCCL on GPUs; gloo in the CPU tests, where
    `backend` is an emulation of the four device operations)."""

    def __init__(self, equation, A, group=None, backend=None, device=None, symbol_pos_local=None, transport="auto"):
        from .linefft import geom, to_position_order

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        P = self.world
        self.Nx, self.Ny, self.Nz = (int(p) for p in equation.domain.points)
        if self.Nx % P or self.Ny % P:
            raise ValueError("Nx and Ny must be divisible by the number of ranks")
        self.nxl, self.C = self.Nx // P, self.Ny // P
        if self.nxl < 2:
            raise ValueError("each slab needs at least 2 planes (halo width)")
        nxl, C, Nx, Ny, Nz = self.nxl, self.C, self.Nx, self.Ny, self.Nz
        self.device = device
        if backend is None:
            from .fused import Ch3dPlan

            plan = Ch3dPlan((nxl, Ny, Nz), equation.domain.dx, equation.kappa, equation._mu_c.descriptor(), equation._mob_c.descriptor())
            backend = _DeviceBackend(plan)
        self.backend = backend
        # local part of A*symbol: [Nx][C][Hz] for this rank's y-chunk (position order along x and y,
        # natural kz = 0..Nz/2 along z)
        Hz = self.Hz = Nz // 2 + 1
        if symbol_pos_local is None:
            import numpy as np

            s = np.asarray(equation.fourier_symbol)
            s = (np.float32(A) * s.real.astype(np.float32)).astype(np.float32)
            s = to_position_order(s, (0, 1))[:, self.rank * C : (self.rank + 1) * C, :Hz]
            symbol_pos_local = torch.from_numpy(np.ascontiguousarray(s)).to(device)
        self.sym = symbol_pos_local
        self.scale = 1.0 / float(Nx * Ny * Nz)
        # line geometries (elements)
        self.g_y = geom(nxl * Hz, Hz, Ny * Hz, 1, Ny, Hz)
        self.g_y_packed = geom(nxl * Hz, Hz, C * Hz, 1, Ny, Hz, chunk=C, hi=nxl * C * Hz)
        self.g_x = geom(C * Hz, C * Hz, 0, 1, Nx, C * Hz)
        self._bufs = None
        big = nxl * Ny * (Nz // 2 + 1) >= (1 << 24)  # complex elements per rank and transpose
        import os as _os

        self._timing = [] if _os.environ.get("PDEOPT_SLAB_TIMING") == "1" else None
        # chunks of the pipelined push transport (1 = transforms and pushes strictly in turn); the chunk boundaries must keep
        # every pushed run a multiple of 16 bytes: even chunk sizes along the local y range
        self._chunks = int(_os.environ.get("PDEOPT_SLAB_CHUNKS", "4" if big else "1"))
        while self._chunks > 1 and (nxl % self._chunks or C % (2 * self._chunks)):
            self._chunks //= 2
        self._side = None
        self._pipe_geoms = None
        if transport not in ("auto", "push", "peer", "nccl"):
            raise ValueError("transport must be 'auto', 'push', 'peer' or 'nccl'")
        on_gpu = isinstance(self.backend, _DeviceBackend) and device is not None and torch.device(device).type == "cuda"
        if transport == "auto":
            # P2P pushes on the GPUs of one box, collectives otherwise (CPU emulation, one rank).  Measured at 512^3
            # (profiles/slab_multigpu_r2_*.log): 2 ranks — push pipelined in 4 chunks 1.48 ms, fused peer stores 1.58, push
            # unpipelined 1.63; 8 ranks — push 0.57, fused peer stores 0.65, NCCL 0.70 (pipelining: no gain, the chunks are
            # launch-bound).  Small slabs on two ranks (half of every buffer stays local): fused peer stores.
            transport = ("push" if (self.world > 2 or big) else "peer") if (on_gpu and self.world > 1) else "nccl"
        self.transport = transport if self.world > 1 else "nccl"
        if self.transport in ("peer", "push"):
            import torch.distributed._symmetric_memory as symm_mem

            n = nxl * Ny * Hz  # complex elements per rank in either packed layout
            grp = group if group is not None else dist.group.WORLD
            self._sym_recv1 = symm_mem.empty(2 * n, dtype=torch.float32, device=device)  # x-line layout [Nx][C][Hz]
            self._sym_recv2 = symm_mem.empty(2 * n, dtype=torch.float32, device=device)  # packed y layout [P][nxl][C][Hz]
            self._h1 = symm_mem.rendezvous(self._sym_recv1, grp)
            self._h2 = symm_mem.rendezvous(self._sym_recv2, grp)
            self._peers1 = list(self._h1.buffer_ptrs)
            self._peers2 = list(self._h2.buffer_ptrs)
            # destination geometries inside one peer's buffer (hi = 0: the chunk index selects the peer)
            self.g_y_to_peers = geom(nxl * Hz, Hz, C * Hz, 1, Ny, Hz, chunk=C, hi=0)
            self.g_x_to_peers = geom(C * Hz, C * Hz, 0, 1, Nx, C * Hz, chunk=nxl, hi=0)
            # halo planes by P2P stores as well: [lo: planes -2, -1 | hi: planes nxl, nxl+1] of THIS rank, written by its
            # two ring neighbours straight into this symmetric buffer (no collective; one symmetric-memory barrier)
            self._sym_halo = symm_mem.empty(4 * Ny * Nz, dtype=torch.float32, device=device)
            self._h0 = symm_mem.rendezvous(self._sym_halo, grp)
            up, down = (self.rank + 1) % P, (self.rank - 1) % P
            self._halo_up_lo = self._h0.get_buffer(up, (2, Ny, Nz), torch.float32, 0)              # my last planes -> up.lo
            self._halo_down_hi = self._h0.get_buffer(down, (2, Ny, Nz), torch.float32, 2 * Ny * Nz)  # my first planes -> down.hi
            self._halo_mine = self._sym_halo.view(2, 2, Ny, Nz)

    def _buffers(self, like):
        if self._bufs is None:
            n = self.nxl * self.Ny * self.Hz
            mk = lambda: torch.empty(n, dtype=torch.complex64, device=like.device)
            pl = (2, self.Ny, self.Nz)
            self._bufs = dict(W=mk(), send=mk(), recv=mk(), lo=torch.empty(pl, dtype=like.dtype, device=like.device),
                              hi=torch.empty(pl, dtype=like.dtype, device=like.device))
        return self._bufs

    def exchange_halos(self, u):
        """halo_lo = last two planes of rank-1, halo_hi = first two planes of rank+1 (periodic ring)."""
        b = self._buffers(u)
        if self.world == 1:
            b["lo"].copy_(u[-2:])
            b["hi"].copy_(u[:2])
            return b["lo"], b["hi"]
        if self.transport in ("peer", "push"):
            # the planes the neighbours need go straight into their halo buffers; the barrier orders the stores before
            # the right-hand side reads them, and the two transpose barriers of the step order those reads before the
            # next step's stores
            if isinstance(self.backend, _DeviceBackend):
                nb = 2 * self.Ny * self.Nz * 4
                self.backend.push_blocks(u[-2:], [self._h0.buffer_ptrs[(self.rank + 1) % self.world]], nb, 0, 0)
                self.backend.push_blocks(u[:2], [self._h0.buffer_ptrs[(self.rank - 1) % self.world]], nb, nb, 0)
            else:
                self._halo_up_lo.copy_(u[-2:])
                self._halo_down_hi.copy_(u[:2])
            self._h0.barrier(channel=2)
            return self._halo_mine[0], self._halo_mine[1]
        # one small all-gather of the four boundary planes of every rank (4 x Ny x Nz floats each);
        # cheap next to the slab transposes and identical on NCCL and gloo
        mine = torch.cat([u[:2], u[-2:]], 0).contiguous()
        allb = torch.empty((self.world * 4,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(allb, mine, group=self.group)
        allb = allb.view((self.world, 4) + tuple(mine.shape[1:]))
        up, down = (self.rank + 1) % self.world, (self.rank - 1) % self.world
        b["lo"].copy_(allb[down, 2:4])
        b["hi"].copy_(allb[up, 0:2])
        return b["lo"], b["hi"]

    def _all_to_all(self, dst, src):
        if self.world == 1:
            dst.copy_(src)
        else:
            dist.all_to_all_single(torch.view_as_real(dst), torch.view_as_real(src), group=self.group)

    def _mark(self, name):
        """PDEOPT_SLAB_TIMING=1: CUDA events between the phases of a step (tools/bench_configs.py prints the averages)."""
        if self._timing is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._timing.append((name, ev))

    def phase_times_ms(self):
        """Average time of every phase over the steps recorded since the last call (needs PDEOPT_SLAB_TIMING=1)."""
        torch.cuda.synchronize()
        acc, cnt = {}, {}
        for (n0, e0), (n1, e1) in zip(self._timing[:-1], self._timing[1:]):
            if n1 != "start":
                acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
                cnt[n1] = cnt.get(n1, 0) + 1
        self._timing = []
        return {k: acc[k] / cnt[k] for k in acc}

    def step(self, u, dt, out=None):
        """u: this rank's slab [nxl, Ny, Nz] float32; returns the slab after one step of length dt."""
        be, b = self.backend, self._buffers(u)
        y1 = out if out is not None else torch.empty_like(u)
        self._mark("start")
        lo, hi = self.exchange_halos(u)
        self._mark("halo")
        f = be.rhs(u, lo, hi)
        self._mark("rhs")
        if not (self.transport == "push" and self._chunks > 1):
            be.fft_r2c(f, b["W"], self.Nz, self.nxl * self.Ny)
            self._mark("r2c_z")
        if self.transport == "peer":
            blk = self.rank * self.nxl * self.C * self.Hz
            r1 = torch.view_as_complex(self._sym_recv1.view(-1, 2))
            r2 = torch.view_as_complex(self._sym_recv2.view(-1, 2))
            # y pass -> peers' x-line buffers; barrier; x pass (fwd * m * inv) -> peers' packed y buffers; barrier
            be.fft_lines_to_peers(b["W"], self.Ny, self.g_y, self._peers1, self.g_y_to_peers, blk)
            self._mark("y_fwd_to_peers")
            self._h1.barrier(channel=0)
            self._mark("barrier1")
            be.fft_lines_to_peers(r1, self.Nx, self.g_x, self._peers2, self.g_x_to_peers, blk, self.sym, self.g_x, dt, self.scale)
            self._mark("x_imex_to_peers")
            self._h2.barrier(channel=1)
            self._mark("barrier2")
            be.fft_lines(r2, b["W"], self.Ny, self.g_y_packed, self.g_y, True, False, 1.0)
            self._mark("y_inv")
            be.fft_c2r_update(b["W"], self.Nz, self.nxl * self.Ny, u, y1, dt)
            self._mark("c2r_z_update")
            return y1
        if self.transport == "push" and self._chunks > 1:
            return self._step_pipelined(u, f, dt, y1)
        if self.transport == "push":
            # transforms write packed local buffers at HBM speed; each transpose is one push kernel (P2P stores, block p ->
            # peer p) followed by a symmetric-memory barrier
            blk_bytes = self.nxl * self.C * self.Hz * 8
            r1 = torch.view_as_complex(self._sym_recv1.view(-1, 2))
            r2 = torch.view_as_complex(self._sym_recv2.view(-1, 2))
            be.fft_lines(b["W"], b["send"], self.Ny, self.g_y, self.g_y_packed, False, False, 1.0)
            self._mark("y_fwd")
            be.push_blocks(b["send"], self._peers1, blk_bytes, self.rank * blk_bytes, self.rank + 1)
            self._mark("push1")
            self._h1.barrier(channel=0)
            self._mark("barrier1")
            be.fft_lines_imex(r1, self.Nx, self.g_x, self.sym, self.g_x, dt, self.scale)
            self._mark("x_imex")
            be.push_blocks(r1, self._peers2, blk_bytes, self.rank * blk_bytes, self.rank + 1)
            self._mark("push2")
            self._h2.barrier(channel=1)
            self._mark("barrier2")
            be.fft_lines(r2, b["W"], self.Ny, self.g_y_packed, self.g_y, True, False, 1.0)
            self._mark("y_inv")
            be.fft_c2r_update(b["W"], self.Nz, self.nxl * self.Ny, u, y1, dt)
            self._mark("c2r_z_update")
            return y1
        be.fft_lines(b["W"], b["send"], self.Ny, self.g_y, self.g_y_packed, False, False, 1.0)
        self._all_to_all(b["recv"], b["send"])
        be.fft_lines_imex(b["recv"], self.Nx, self.g_x, self.sym, self.g_x, dt, self.scale)
        self._all_to_all(b["send"], b["recv"])
        be.fft_lines(b["send"], b["W"], self.Ny, self.g_y_packed, self.g_y, True, False, 1.0)
        be.fft_c2r_update(b["W"], self.Nz, self.nxl * self.Ny, u, y1, dt)
        return y1

    def _step_pipelined(self, u, f, dt, y1):
        """transport="push" with the transposes pipelined against the transforms (PDEOPT_SLAB_CHUNKS chunks, two side
        streams): transpose #1 is cut along the slab's planes — z transform, y transform and push of chunk q+1 run while
        chunk q is on the wire; transpose #2 is cut along the local y range — x transform (fwd * m * inv) of chunk q+1 under
        the push of chunk q.  One symmetric-memory barrier per transpose, as before; same arithmetic, bit for bit."""
        from .linefft import geom

        be, b = self.backend, self._buffers(u)
        nxl, C, Hz, Ny, Nx, Nz, Q = self.nxl, self.C, self.Hz, self.Ny, self.Nx, self.Nz, self._chunks
        blk = nxl * C * Hz  # complex elements per destination block
        main = torch.cuda.current_stream(u.device)
        if self._side is None:
            self._side = [torch.cuda.Stream(u.device), torch.cuda.Stream(u.device)]
        r1 = torch.view_as_complex(self._sym_recv1.view(-1, 2))
        r2 = torch.view_as_complex(self._sym_recv2.view(-1, 2))
        W, send, symf = b["W"], b["send"], self.sym.view(-1)
        if self._pipe_geoms is None:
            self._pipe_geoms = ([(geom((nxl // Q) * Hz, Hz, Ny * Hz, 1, Ny, Hz), geom((nxl // Q) * Hz, Hz, C * Hz, 1, Ny, Hz, chunk=C, hi=blk))
                                 for _ in range(Q)],
                                [geom((C // Q) * Hz, (C // Q) * Hz, 0, 1, Nx, C * Hz) for _ in range(Q)])
        ev = main.record_event()
        for q in range(Q):
            x0, x1 = q * nxl // Q, (q + 1) * nxl // Q
            st = self._side[q % 2]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                be.fft_r2c(f[x0:x1], W[x0 * Ny * Hz:], Nz, (x1 - x0) * Ny)
                be.fft_lines(W[x0 * Ny * Hz:], send[x0 * C * Hz:], Ny, self._pipe_geoms[0][q][0], self._pipe_geoms[0][q][1], False, False, 1.0)
                be.push_rows(send[x0 * C * Hz:], self._peers1, blk * 8, (self.rank * blk + x0 * C * Hz) * 8, 1, (x1 - x0) * C * Hz * 8, 0,
                             self.rank + 1)
        for st in self._side:
            main.wait_stream(st)
        self._mark("z_y_push1")
        self._h1.barrier(channel=0)
        self._mark("barrier1")
        ev = main.record_event()
        for q in range(Q):
            c0, c1 = q * C // Q, (q + 1) * C // Q
            st = self._side[q % 2]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                g = self._pipe_geoms[1][q]
                be.fft_lines_imex(r1[c0 * Hz:], Nx, g, symf[c0 * Hz:], g, dt, self.scale)
                be.push_rows(r1[c0 * Hz:], self._peers2, blk * 8, (self.rank * blk + c0 * Hz) * 8, nxl, (c1 - c0) * Hz * 8, C * Hz * 8,
                             self.rank + 1)
        for st in self._side:
            main.wait_stream(st)
        self._mark("x_imex_push2")
        self._h2.barrier(channel=1)
        self._mark("barrier2")
        be.fft_lines(r2, W, Ny, self.g_y_packed, self.g_y, True, False, 1.0)
        self._mark("y_inv")
        be.fft_c2r_update(W, Nz, nxl * Ny, u, y1, dt)
        self._mark("c2r_z_update")
        return y1

    def rollout(self, u, dts):
        y = u
        for dt in dts:
            y = self.step(y, float(dt))
        return y
