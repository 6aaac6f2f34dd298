"""Register-FFT building blocks (pde_opt_b200/csrc/regfft.cuh) compiled for the host and
checked against numpy.fft: forward DIF (bit-reversed out) and inverse DIT (bit-reversed in)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("regfft") / "regfft_host.so"
    src = os.path.join(HERE, "host", "regfft_host.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)])
    return ctypes.CDLL(str(out))


def brev(k, bits):
    return int(format(k, f"0{bits}b")[::-1], 2)


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32])
def test_dif_forward_and_dit_inverse(lib, n):
    rng = np.random.default_rng(n)
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)
    buf = np.ascontiguousarray(x.view(np.float32).copy())
    lib.dif_fwd(n, buf.ctypes.data_as(ctypes.c_void_p))
    got = buf.view(np.complex64)
    bits = n.bit_length() - 1
    perm = np.array([brev(p, bits) for p in range(n)])
    ref = np.fft.fft(x.astype(np.complex128))
    np.testing.assert_allclose(got, ref[perm], rtol=0, atol=2e-6 * np.abs(ref).max())
    # inverse DIT consumes the bit-reversed spectrum and returns n * x in natural order
    lib.dit_inv(n, buf.ctypes.data_as(ctypes.c_void_p))
    np.testing.assert_allclose(buf.view(np.complex64) / n, x, rtol=0, atol=4e-6 * np.abs(x).max())


@pytest.mark.parametrize("n", [4, 8, 16, 32])
def test_dif_inverse_and_dit_forward(lib, n):
    rng = np.random.default_rng(100 + n)
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)
    bits = n.bit_length() - 1
    perm = np.array([brev(p, bits) for p in range(n)])
    buf = np.ascontiguousarray(x.view(np.float32).copy())
    lib.dif_inv(n, buf.ctypes.data_as(ctypes.c_void_p))
    ref = np.fft.ifft(x.astype(np.complex128)) * n
    np.testing.assert_allclose(buf.view(np.complex64), ref[perm], rtol=0, atol=2e-6 * np.abs(ref).max())
    buf = np.ascontiguousarray(x[perm].view(np.float32).copy())
    lib.dit_fwd(n, buf.ctypes.data_as(ctypes.c_void_p))
    ref = np.fft.fft(x.astype(np.complex128))
    np.testing.assert_allclose(buf.view(np.complex64), ref, rtol=0, atol=2e-6 * np.abs(ref).max())


@pytest.mark.parametrize("n", [4, 8, 16, 32])
def test_ditf_forward_and_inverse(lib, n):
    """DitF: every non-trivial twiddle in the six-slot FMA form (bit-reversed in, natural out)."""
    rng = np.random.default_rng(200 + n)
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)
    bits = n.bit_length() - 1
    perm = np.array([brev(p, bits) for p in range(n)])
    buf = np.ascontiguousarray(x[perm].view(np.float32).copy())
    lib.ditf_fwd(n, buf.ctypes.data_as(ctypes.c_void_p))
    ref = np.fft.fft(x.astype(np.complex128))
    np.testing.assert_allclose(buf.view(np.complex64), ref, rtol=0, atol=3e-6 * np.abs(ref).max())
    buf = np.ascontiguousarray(x[perm].view(np.float32).copy())
    lib.ditf_inv(n, buf.ctypes.data_as(ctypes.c_void_p))
    ref = np.fft.ifft(x.astype(np.complex128)) * n
    np.testing.assert_allclose(buf.view(np.complex64), ref, rtol=0, atol=3e-6 * np.abs(ref).max())
