"""On-disk formats either side of the dose path: `.dat` dose kernels in, NIfTI-1 dose maps out."""
from .kernel_dat import load_kernel, save_kernel
from .nifti import load_dose_map, save_dose_map

__all__ = ["load_kernel", "save_kernel", "load_dose_map", "save_dose_map"]
