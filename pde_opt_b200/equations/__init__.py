from .advection_diffusion import AdvectionDiffusion2D
from .base_eq import BaseEquation, TimeSplittingEquation
from .gross_pitaevskii import GPE2DTSControl
from .phase_field import AllenCahn2DPeriodic, CahnHilliard2DPeriodic, CahnHilliard3DPeriodic

__all__ = ["BaseEquation", "TimeSplittingEquation", "CahnHilliard2DPeriodic", "CahnHilliard3DPeriodic", "AllenCahn2DPeriodic", "GPE2DTSControl", "AdvectionDiffusion2D"]
