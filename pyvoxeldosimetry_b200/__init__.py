"""pyvoxeldosimetry_b200 - B200-native (sm_100a) kernel-convolution dose path behind the
PyVoxelDosimetry calculator API.  Importing the package is cheap; the CUDA library is loaded on first
use and there is NO CPU fallback (a missing libpvdose.so or CUDA device raises)."""
__version__ = "0.1.0"

from .core import (ActivitySampler, DoseCalculationResult, DoseCalculator, DosimetryCalculator,
                   KernelConvolutionCalculator)
from .data.dose_kernels import KernelFactory
from .time_integration import TimeCurveFitting

__all__ = ["DosimetryCalculator", "KernelConvolutionCalculator", "ActivitySampler", "DoseCalculator",
           "DoseCalculationResult", "KernelFactory", "TimeCurveFitting"]
