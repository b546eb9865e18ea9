"""Core of the drop-in: same names as pyvoxeldosimetry.core (reference core/__init__.py:7-23), kernel path only."""
from .activity_sampler import ActivitySampler
from .dose_calculator import DoseCalculationResult, DoseCalculator
from .dosimetry_base import DosimetryCalculator
from .kernel_convolution import KernelConvolutionCalculator, trapezoid_weights

__all__ = ["DosimetryCalculator", "KernelConvolutionCalculator", "ActivitySampler", "DoseCalculator",
           "DoseCalculationResult", "trapezoid_weights"]
