"""The real-field spectral filter of the 128x128 kernel (pde_opt_b200/csrc/rfft128.cuh) executed on
the host: 256 emulated threads per barrier phase on a buffer standing in for shared memory, checked
against numpy.fft (the reference's jnp.fft.fftn / ifftn pair, solvers.py:62-63)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
N = 128


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("rfft128") / "rfft128_host.so"
    src = os.path.join(HERE, "host", "rfft128_host.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)])
    return ctypes.CDLL(str(out))


def folded_symbol(h=0.01, kappa=0.002, A=0.5):
    k = np.fft.fftfreq(N, h)
    k2 = (2 * np.pi) ** 2 * (k[:, None] ** 2 + k[None, :] ** 2)
    sym = A * kappa * k2 * k2
    return sym, np.ascontiguousarray(sym[:65, :65].astype(np.float32))


def run(lib, f, tab, dt):
    g = np.empty((N, N), np.float32)
    rt = np.empty((N, N), np.float32)
    P = ctypes.c_void_p
    lib.rfft128_filter(f.ctypes.data_as(P), tab.ctypes.data_as(P), ctypes.c_float(dt), g.ctypes.data_as(P), rt.ctypes.data_as(P))
    return g, rt


@pytest.mark.parametrize("dt", [1e-6, 1e-4, 0.0])
def test_filter_matches_numpy(lib, dt):
    rng = np.random.default_rng(3)
    f = rng.normal(size=(N, N)).astype(np.float32)
    sym, tab = folded_symbol()
    g, rt = run(lib, f, tab, dt)
    ref = np.fft.ifft2(np.fft.fft2(f.astype(np.float64)) / (1.0 + dt * sym)).real
    err = np.linalg.norm(g - ref) / np.linalg.norm(ref)
    assert err < 2e-6, err
    # scatter_nat is the inverse of gather_nat
    np.testing.assert_array_equal(rt, g)


def test_single_modes(lib):
    """Every class of wavenumbers (self-conjugate rows / columns, Nyquist lines, generic)."""
    sym, tab = folded_symbol()
    x = np.arange(N)
    for kr, kc in [(0, 0), (64, 0), (0, 64), (64, 64), (0, 32), (8, 0), (8, 32), (16, 5), (3, 0), (3, 32), (3, 64),
                   (5, 7), (120, 99), (64, 17), (77, 64), (40, 96)]:
        f = (np.cos(2 * np.pi * (kr * x[:, None] + kc * x[None, :]) / N)
             + 0.5 * np.sin(2 * np.pi * (kr * x[:, None] + kc * x[None, :]) / N)).astype(np.float32)
        g, _ = run(lib, f, tab, 1e-6)
        ref = np.fft.ifft2(np.fft.fft2(f.astype(np.float64)) / (1.0 + 1e-6 * sym)).real
        assert np.abs(g - ref).max() < 5e-6 * max(1.0, np.abs(ref).max()), (kr, kc)
