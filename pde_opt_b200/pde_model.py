"""PDEModel.solve — mirror of pde_opt/pde_model.py:15-136 for the stepping path.

`train` / `optimize` / `residuals` (pde_model.py:138-551) sit on top of the adjoint and are
SURVEY 8(f) "next" rows."""
from typing import Any, Dict

import numpy as np
import torch

from .schedule import constant_step_times
from .solvers import ODETerm
from .utils import check_equation_solver_compatibility, prepare_solver_params


class PDEModel:
    def __init__(self, equation_type, domain, solver_type):
        self.equation_type = equation_type
        self.domain = domain
        self.solver_type = solver_type
        check_equation_solver_compatibility(self.solver_type, self.equation_type)  # pde_model.py:66

    def solve(self, parameters: Dict[str, Any], y0, ts, solver_parameters: Dict[str, Any] = {}, adjoint=None,
              dt0=0.000001, max_steps=1000000, stepsize_controller=None):
        """Integrate from ts[0] to ts[-1] with constant step dt0 and return the solution at `ts`
        (shape (len(ts), *y0.shape)), linearly interpolated inside the step that brackets each
        save time, as diffrax's SaveAt(ts=ts) does with LocalLinearInterpolation
        (pde_model.py:120-136).  `adjoint` / `stepsize_controller` are accepted for signature
        compatibility; only the constant-step forward solve is implemented here."""
        if stepsize_controller is not None and type(stepsize_controller).__name__ != "ConstantStepSize":
            raise NotImplementedError("only ConstantStepSize is implemented on the fused path")
        equation = self.equation_type(domain=self.domain, **parameters)  # :110
        solver = self.solver_type(**prepare_solver_params(self.solver_type, solver_parameters, equation))  # :112-117
        terms = ODETerm(equation)
        ts = np.asarray([float(t) for t in ts], dtype=np.float32)
        times = constant_step_times(ts[0], ts[-1], dt0, np.float32, max_steps)
        truncated = times[-1] < ts[-1]
        y = y0 if torch.is_tensor(y0) else torch.as_tensor(np.asarray(y0, dtype=np.float32))
        if not y.is_cuda:
            y = y.cuda()
        y = y.to(torch.float32).contiguous()
        out = torch.empty((len(ts),) + tuple(y.shape), dtype=torch.float32, device=y.device)
        i_cur = 0
        for si, s in enumerate(ts):
            j = int(np.searchsorted(times, s, side="left"))
            if j >= len(times):
                out[si:] = float("inf")  # diffrax (throw=False) leaves unreached save slots at inf
                break
            if j == 0 or times[j] == s:
                if j > i_cur:
                    y = solver.rollout(terms, times[i_cur : j + 1], y)
                    i_cur = j
                out[si] = y
                continue
            if j - 1 > i_cur:
                y = solver.rollout(terms, times[i_cur:j], y)
                i_cur = j - 1
            y_b = solver.rollout(terms, times[j - 1 : j + 1], y)
            w = np.float32((s - times[j - 1]) / (times[j] - times[j - 1]))
            out[si] = torch.lerp(y, y_b, float(w))
            y, i_cur = y_b, j
        del truncated
        return out
