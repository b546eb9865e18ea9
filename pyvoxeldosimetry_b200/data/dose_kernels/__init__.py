from .generators import BaseKernelGenerator, F18KernelGenerator, Ga68KernelGenerator, Lu177KernelGenerator, Y90KernelGenerator
from .kernel_factory import KernelFactory

__all__ = ["BaseKernelGenerator", "Y90KernelGenerator", "Lu177KernelGenerator", "Ga68KernelGenerator", "F18KernelGenerator", "KernelFactory"]
