"""Geometry for the smoothed-boundary equations.  Mirror of pde_opt/numerics/shapes.py:20-203: a binary mask is
smoothed into a level set psi in (0, 1] by an Allen-Cahn flow with curvature control (setup-time host work, one field),
and the graph-Laplacian eigenmodes of the mask are available as a shape basis.

The reference integrates the smoothing flow with diffrax Tsit5 + PIDController(rtol=1e-4, atol=1e-6) (shapes.py:67-77).
diffrax is not installable here; the same flow is integrated with SciPy's adaptive Dormand-Prince pair (RK45) at the
same tolerances and first step — same order, same error control, different tableau: the smoothed field agrees with the
reference's to the integration tolerance, not bit for bit."""
import dataclasses
from typing import Optional, Tuple

import numpy as np


def _gx(a, h):  # utils/derivatives.py:69-71
    return 0.5 * (np.roll(a, -1, 0) - np.roll(a, 1, 0)) / h


def _gy(a, h):  # :74-76
    return 0.5 * (np.roll(a, -1, 1) - np.roll(a, 1, 1)) / h


def _gxx(a, h):  # :84-86
    return (np.roll(a, -1, 0) - 2 * a + np.roll(a, 1, 0)) / h**2


def _gyy(a, h):  # :89-91
    return (np.roll(a, -1, 1) - 2 * a + np.roll(a, 1, 1)) / h**2


def _gxy(a, hx, hy):  # :99-106
    return (np.roll(np.roll(a, -1, 0), -1, 1) + np.roll(np.roll(a, 1, 0), 1, 1) - np.roll(np.roll(a, -1, 0), 1, 1)
            - np.roll(np.roll(a, 1, 0), -1, 1)) / (4.0 * hx * hy)


@dataclasses.dataclass
class Shape:
    binary: np.ndarray
    dx: Optional[Tuple[float, float]] = (1.0, 1.0)
    smooth_epsilon: float = 1.0
    smooth_curvature: float = 0.0
    smooth_dt: float = 0.1
    smooth_tf: float = 1.0

    def __post_init__(self):
        self.binary = np.asarray(self.binary, dtype=np.float64)
        s = self.smooth_shape()
        s = np.where(s < 0.001, 0.001, s)  # shapes.py:37-38
        self.smooth = np.where(s > 0.99, 1.0, s)

    def flow_rhs(self, u):
        """shapes.py:43-64: 2 (c lap u + (1 - c) u_nn) - W'(u) / eps with the double-well 18/eps u (1-u)(1-2u)."""
        hx, hy = self.dx
        gx, gy = _gx(u, hx), _gy(u, hy)
        gxx, gyy, gxy = _gxx(u, hx), _gyy(u, hy), _gxy(u, hx, hy)
        gn = gx**2 + gy**2
        gn = np.where(gn < 1e-7, 1.0, gn)
        norm_lap = (gxx * gx**2 + 2.0 * gxy * gx * gy + gyy * gy**2) / gn
        pot = 18.0 / self.smooth_epsilon * u * (1.0 - u) * (1.0 - 2.0 * u)
        return 2.0 * (self.smooth_curvature * (gxx + gyy) + (1.0 - self.smooth_curvature) * norm_lap) - pot / self.smooth_epsilon

    def smooth_shape(self):
        from scipy.integrate import solve_ivp

        shp = self.binary.shape
        sol = solve_ivp(lambda t, y: self.flow_rhs(y.reshape(shp)).ravel(), (0.0, float(self.smooth_tf)), self.binary.ravel(),
                        method="RK45", rtol=1e-4, atol=1e-6, first_step=min(float(self.smooth_dt), float(self.smooth_tf)))
        return sol.y[:, -1].reshape(shp)

    def laplacian_from_mask(self, periodic: bool = False):
        """Unnormalised 4-neighbour graph Laplacian of the pixels where the mask is 1 (shapes.py:81-146): returns the
        CSR matrix and the (H, W) array of node indices (-1 outside the mask)."""
        from scipy.sparse import coo_matrix, csr_matrix

        mask = self.binary > 0
        ids = np.full(mask.shape, -1, dtype=np.int64)
        n = int(mask.sum())
        ids[mask] = np.arange(n)
        if n == 0:
            return csr_matrix((0, 0)), ids
        us, vs = [], []
        for axis in (1, 0):  # right neighbours, then down neighbours: every undirected edge once
            if periodic:
                nb_mask, nb_ids = np.roll(mask, -1, axis), np.roll(ids, -1, axis)
                both = mask & nb_mask
                us.append(ids[both]); vs.append(nb_ids[both])
            else:
                a = [slice(None)] * 2; b = [slice(None)] * 2
                a[axis], b[axis] = slice(0, -1), slice(1, None)
                both = mask[tuple(a)] & mask[tuple(b)]
                us.append(ids[tuple(a)][both]); vs.append(ids[tuple(b)][both])
        u, v = np.concatenate(us), np.concatenate(vs)
        deg = np.bincount(np.concatenate([u, v]), minlength=n).astype(np.float64)
        rows = np.concatenate([u, v, np.arange(n)])
        cols = np.concatenate([v, u, np.arange(n)])
        data = np.concatenate([-np.ones(2 * len(u)), deg])
        return coo_matrix((data, (rows, cols)), shape=(n, n)).tocsr(), ids

    def get_shape_modes(self, N: Optional[int] = None):
        """First N eigenvectors of the mask's graph Laplacian scattered back onto the grid (shapes.py:148-203):
        sets `shape_basis` [H, W, N] and `shape_basis_evals`."""
        import scipy.sparse.linalg

        L, ids = self.laplacian_from_mask()
        n = L.shape[0]
        if (L != L.T).nnz != 0:
            raise ValueError("Laplacian matrix is not symmetric")
        sigma = max(float(L.diagonal().mean()) if n > 0 else 1.0, 1.0) * 1e-8
        evals, evecs = scipy.sparse.linalg.eigsh(L, k=N, which="LM", sigma=sigma, tol=1e-8)
        out = np.zeros(self.binary.shape + (N,))
        valid = ids >= 0
        for i in range(N):
            out[valid, i] = evecs[ids[valid], i]
        self.shape_basis, self.shape_basis_evals = out, evals
