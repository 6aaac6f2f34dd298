"""The cluster-distributed 256x256 complex FFT (pde_opt_b200/csrc/cfft256.cuh) executed on the host: eight
emulated CTAs x 256 threads (and the 4 x 512 variant), remote stores as writes into the peers' buffers, against numpy.fft."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
N = 256


@pytest.fixture(scope="module", params=[8, 4])
def lib(request, tmp_path_factory):
    out = tmp_path_factory.mktemp(f"cfft256_{request.param}") / "cfft256_host.so"
    src = os.path.join(HERE, "host", "cfft256_host.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-DPDEOPT_CF_CTAS={request.param}", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)])
    return ctypes.CDLL(str(out))


def test_forward_multiply_inverse(lib):
    rng = np.random.default_rng(0)
    z = (rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))).astype(np.complex64)
    k = np.fft.fftfreq(N, 0.1)
    a_term = 0.5j * (-(2 * np.pi) ** 2) * (k[:, None] ** 2 + k[None, :] ** 2)
    mult = np.exp(a_term * 0.5 * (1e-3 - 2e-4j)).astype(np.complex64)
    inp = np.ascontiguousarray(z.view(np.float32).reshape(N, N, 2))
    m = np.ascontiguousarray(mult.view(np.float32).reshape(N, N, 2))
    out = np.empty_like(inp)
    spec = np.empty_like(inp)
    P = ctypes.c_void_p
    lib.cfft256_roundtrip(inp.ctypes.data_as(P), m.ctypes.data_as(P), out.ctypes.data_as(P), spec.ctypes.data_as(P))
    Z = np.fft.fft2(z.astype(np.complex128))
    got_spec = spec.reshape(N, N * 2).view(np.complex64)
    assert np.abs(got_spec - Z).max() / np.abs(Z).max() < 3e-6
    ref = np.fft.ifft2(Z * mult.astype(np.complex128))
    got = out.reshape(N, N * 2).view(np.complex64) / (N * N)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 5e-6


def test_identity(lib):
    rng = np.random.default_rng(1)
    z = (rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))).astype(np.complex64)
    inp = np.ascontiguousarray(z.view(np.float32).reshape(N, N, 2))
    out = np.empty_like(inp)
    lib.cfft256_roundtrip(inp.ctypes.data_as(ctypes.c_void_p), None, out.ctypes.data_as(ctypes.c_void_p), None)
    got = out.reshape(N, N * 2).view(np.complex64) / (N * N)
    assert np.abs(got - z).max() < 5e-6 * np.abs(z).max()


@pytest.mark.parametrize("ctas", [8, 4])
def test_layout_is_bank_conflict_free(ctas):
    """Every access pattern of the passes hits 16 distinct 8-byte banks per half-warp (16 consecutive lines), and a
    warp's transposed store is 32 consecutive positions of one destination line (256 contiguous bytes) — for the
    8-CTA x 256-thread cluster (32 lines per CTA, the default) and the 4-CTA x 512-thread one."""
    lines_per_cta = 256 // ctas
    slot = lambda line, pos: line * 256 + (pos ^ (line & 15))
    e1 = lambda line, k1, j: line * 256 + 32 * j + (k1 ^ (line & 15))
    ok = lambda s: len({int(v) % 16 for v in s}) == 16
    for l0 in range(0, lines_per_cta, 16):
        lines = range(l0, l0 + 16)
        for j in range(8):
            for n1 in range(32):
                assert ok([slot(l, 8 * n1 + j) for l in lines])                      # spatial load / store
                assert ok([e1(l, n1, j) for l in lines])                             # E1, own j
            for a in range(4):
                for jj in range(8):
                    assert ok([e1(l, j + 8 * a, jj) for l in lines])                 # E1, the other threads' values
                for k0 in range(8):
                    assert ok([slot(l, j + 8 * a + 32 * k0) for l in lines])         # frequency load
    for q in range(ctas):
        for k in range(lines_per_cta):
            for l0 in range(0, lines_per_cta, 32):
                s = sorted(slot(k, lines_per_cta * q + l) for l in range(l0, l0 + 32))
                assert s == list(range(s[0], s[0] + 32)) and s[0] % 16 == 0          # transposed store: contiguous
