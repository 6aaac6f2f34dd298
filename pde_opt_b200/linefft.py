"""Torch-facing wrapper of the line-FFT engine (pdeopt_fft_lines*, include/pdeopt_b200.h).

The engine leaves spectra in digit-reversed *position order* along each transformed axis (so the
steppers never reorder anything); `fftn` / `ifftn` below add the permutation to natural `fftfreq`
order and are what the equations expose as `eq.fft` / `eq.ifft` (the reference stores
`jnp.fft.fftn / ifftn` there: cahn_hilliard.py:72-73, gross_pitaevskii.py:58-59)."""
import ctypes
import functools

import numpy as np
import torch

from . import _lib


@functools.lru_cache(maxsize=None)
def pos_to_freq(n):
    """freq[p] = frequency index stored at position p after a forward transform of length n."""
    lib = _lib.load()
    out = np.array([lib.pdeopt_fft_pos_to_freq(n, p) for p in range(n)], dtype=np.int64)
    if (out < 0).any():
        raise ValueError(f"line FFT length must be a power of two in [8, 512], got {n}")
    return out


@functools.lru_cache(maxsize=None)
def freq_to_pos(n):
    f = pos_to_freq(n)
    inv = np.empty_like(f)
    inv[f] = np.arange(n)
    return inv


def geom(n_lines, n_inner, outer, inner, n, lo, chunk=None, hi=0):
    g = _lib.LineGeom()
    g.n_lines, g.n_inner, g.outer, g.inner = int(n_lines), int(n_inner), int(outer), int(inner)
    g.chunk = int(chunk if chunk is not None else n)
    g.hi, g.lo = int(hi), int(lo)
    return g


def axis_geom(shape, axis):
    """Geometry of the lines along `axis` of a C-contiguous array of `shape` (leading axes batch)."""
    shape = tuple(int(s) for s in shape)
    n = shape[axis]
    inner = int(np.prod(shape[axis + 1 :], dtype=np.int64))
    outer_count = int(np.prod(shape[:axis], dtype=np.int64))
    return geom(outer_count * inner, inner, n * inner, 1 if inner > 1 else 0, n, inner)


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def fft_lines(x, axis, inverse=False, scale=1.0, out=None):
    """Transform along `axis` of a C-contiguous CUDA tensor (complex64, or float32 for a forward
    transform of real data).  Output complex64 in position order (forward) / natural order (inverse)."""
    assert x.is_cuda and x.is_contiguous()
    in_real = x.dtype == torch.float32
    assert in_real or x.dtype == torch.complex64
    axis = axis % x.dim()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.complex64, device=x.device)
    g = axis_geom(x.shape, axis)
    st = _lib.load().pdeopt_fft_lines(
        ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), x.shape[axis], ctypes.byref(g), ctypes.byref(g),
        1 if inverse else 0, 1 if in_real else 0, float(scale), _stream(x),
    )
    _lib.check(st)
    return out


def _perm(n, device, inverse):
    idx = freq_to_pos(n) if not inverse else pos_to_freq(n)
    return torch.as_tensor(idx, device=device)


def fftn(x, dims=None):
    """Unnormalised forward DFT over `dims` (default: all) in natural fftfreq order; CUDA complex64
    (or float32) in, complex64 out.  Same convention as jnp.fft.fftn."""
    dims = tuple(range(x.dim())) if dims is None else tuple(d % x.dim() for d in dims)
    y = x.contiguous()
    for d in reversed(dims):
        y = fft_lines(y, d)
    for d in dims:
        y = y.index_select(d, _perm(y.shape[d], y.device, False))
    return y


def ifftn(x, dims=None):
    """Inverse of :func:`fftn` (normalised by 1/prod(n)), same convention as jnp.fft.ifftn."""
    dims = tuple(range(x.dim())) if dims is None else tuple(d % x.dim() for d in dims)
    y = x.to(torch.complex64)
    for d in dims:
        y = y.index_select(d, _perm(y.shape[d], y.device, True))
    y = y.contiguous()
    for d in dims:
        y = fft_lines(y, d, inverse=True, scale=1.0 / y.shape[d], out=y)
    return y


def to_position_order(a, axes):
    """Permute a natural-order (fftfreq) NumPy/torch array to the engine's position order along `axes`."""
    for ax in axes:
        idx = pos_to_freq(a.shape[ax])
        a = a.take(idx, axis=ax) if isinstance(a, np.ndarray) else a.index_select(ax, torch.as_tensor(idx, device=a.device))
    return a
