"""Step-size controllers accepted by PDEModel.solve (pde_model.py:120-134 passes them to diffeqsolve).

The reference's drivers default to `diffrax.ConstantStepSize()`; the semi-implicit solver also
returns `y_error = y1 - (y0 + dt f0)` (solvers.py:61-65) so that an adaptive controller can be used.
`PIDController` restates the published algorithm of diffrax's controller of the same name (diffrax is
a third-party dependency of the reference, absent from this image: unpinned restatement, see
DESIGN.md section 2):

    scaled_error = rms( y_error / (atol + rtol * max(|y0|, |y1|)) )
    keep_step    = scaled_error < 1
    factor       = clip( safety * e_n^{-b1} e_{n-1}^{-b2} e_{n-2}^{-b3},  [1 if kept else factormin, factormax] )
    b1 = (icoeff + pcoeff + dcoeff) / k,  b2 = -(pcoeff + 2 dcoeff) / k,  b3 = dcoeff / k,  k = error order
    dt_next      = dt * factor      (the error history advances on accepted steps only)

The step itself runs on the GPU; the controller needs one scalar per step back on the host.
"""
import math


class ConstantStepSize:
    """diffrax.ConstantStepSize(): every step has length dt0 (the last one is clipped to t1)."""


class PIDController:
    def __init__(self, rtol, atol, pcoeff=0.0, icoeff=1.0, dcoeff=0.0, dtmin=None, dtmax=None, force_dtmin=True,
                 factormin=0.2, factormax=10.0, safety=0.9, error_order=None):
        self.rtol, self.atol = float(rtol), float(atol)
        self.pcoeff, self.icoeff, self.dcoeff = float(pcoeff), float(icoeff), float(dcoeff)
        self.dtmin, self.dtmax, self.force_dtmin = dtmin, dtmax, force_dtmin
        self.factormin, self.factormax, self.safety = float(factormin), float(factormax), float(safety)
        self.error_order = error_order

    def init_state(self):
        return (1.0, 1.0)  # inverse scaled errors of the two previous accepted steps

    def scaled_error(self, y0, y1, y_error):
        """rms norm of the scaled error: torch tensors (any device) -> python float (one host sync)."""
        import torch

        scale = self.atol + torch.maximum(y0.abs(), y1.abs()) * self.rtol
        return float(torch.sqrt(torch.mean((y_error / scale).double() ** 2)))

    def adapt(self, dt, scaled_error, solver_order, state):
        """-> (keep_step, next_dt, state)"""
        k = float(self.error_order if self.error_order is not None else solver_order)
        prev, prev_prev = state
        keep = scaled_error < 1.0
        at_dtmin = self.dtmin is not None and dt <= self.dtmin
        if at_dtmin and self.force_dtmin:
            keep = True
        inv = 1.0 / scaled_error if (scaled_error > 0.0 and math.isfinite(scaled_error)) else (1.0 if scaled_error == 0.0 else 0.0)
        b1 = (self.icoeff + self.pcoeff + self.dcoeff) / k
        b2 = -(self.pcoeff + 2.0 * self.dcoeff) / k
        b3 = self.dcoeff / k
        f1 = 1.0 if b1 == 0.0 else inv**b1
        f2 = 1.0 if b2 == 0.0 else prev**b2
        f3 = 1.0 if b3 == 0.0 else prev_prev**b3
        lo = 1.0 if keep else self.factormin
        factor = min(max(self.safety * f1 * f2 * f3, lo), self.factormax)
        nxt = dt * factor
        if self.dtmin is not None:
            nxt = max(nxt, self.dtmin)
        if self.dtmax is not None:
            nxt = min(nxt, self.dtmax)
        if keep:
            state = (inv, prev)
        return keep, nxt, state
