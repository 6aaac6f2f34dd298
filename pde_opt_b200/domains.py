"""Simulation domain.  Mirrors pde_opt/numerics/domains.py:16-64 (same fields, same methods,
NumPy arrays instead of jax arrays; the grids are setup-time host data)."""
import dataclasses
from typing import Optional, Tuple

import numpy as np


@dataclasses.dataclass
class Domain:
    points: Tuple[int, ...]
    box: Tuple[Tuple[float, float], ...]
    units: str
    geometry: Optional[object] = None

    def __post_init__(self):
        # domains.py:29-34
        self.dx = tuple((hi - lo) / n for (lo, hi), n in zip(self.box, self.points))
        self.L = tuple((hi - lo) for (lo, hi) in self.box)

    def axes(self):
        return tuple(  # domains.py:36-42
            np.linspace(lo + h / 2, hi - h / 2, num=n) for (lo, hi), n, h in zip(self.box, self.points, self.dx)
        )

    def fft_axes(self):
        return tuple(np.fft.fftfreq(n, h) for n, h in zip(self.points, self.dx))  # domains.py:44-47

    def rfft_axes(self):
        return tuple(np.fft.rfftfreq(n, h) for n, h in zip(self.points, self.dx))  # domains.py:49-52

    def mesh(self):
        return tuple(np.meshgrid(*self.axes(), indexing="ij"))

    def fft_mesh(self):
        return tuple(np.meshgrid(*self.fft_axes(), indexing="ij"))

    def rfft_mesh(self):
        return tuple(np.meshgrid(*self.rfft_axes(), indexing="ij"))

    def __str__(self):
        return f"Domain with bounds {self.box} with units of {self.units} and {self.points} collocation points."
