from .advection_diffusion import AdvectionDiffusion2D
from .base_eq import BaseEquation, TimeSplittingEquation
from .gross_pitaevskii import GPE2DTSControl
from .phase_field import AllenCahn2DPeriodic, CahnHilliard2DPeriodic, CahnHilliard3DPeriodic
from .smoothed_boundary import AllenCahn2DSmoothedBoundary, CahnHilliard2DSmoothedBoundary

__all__ = ["BaseEquation", "TimeSplittingEquation", "CahnHilliard2DPeriodic", "CahnHilliard3DPeriodic", "AllenCahn2DPeriodic", "GPE2DTSControl", "AdvectionDiffusion2D",
           "CahnHilliard2DSmoothedBoundary", "AllenCahn2DSmoothedBoundary"]
