"""CT Hounsfield units -> mass density for the voxel-wise density correction (A9, spec-defined here:
the reference accepts `tissue_densities` and ignores it, core/dose_calculator.py:90).

Knots follow the material table the reference ships for GATE (core/gate/data/HU_to_material.txt:5-20
with the GateMaterials.db densities: Air 0.00129, Lung 0.26, Adipose 0.92, Water 1.0, Muscle 1.05,
SpineBone 1.42, RibBone 1.92 g/cm3); piecewise linear, clamped at both ends.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import engine

HU_KNOTS = ((-1000.0, 0.00129), (-700.0, 0.26), (-100.0, 0.92), (0.0, 1.0), (40.0, 1.05), (350.0, 1.42),
            (1200.0, 1.92), (3000.0, 2.90))


def hu_to_density(hu, knots=HU_KNOTS, device=None):
    """int16 / float32 HU volume -> float32 density (g/cm3).  Host in -> host out, CUDA in -> CUDA out."""
    dev = engine.require_cuda(device)
    on_dev = isinstance(hu, torch.Tensor) and hu.is_cuda
    if isinstance(hu, torch.Tensor):
        t = hu.to(dev)
    else:
        a = np.ascontiguousarray(hu)
        if a.dtype != np.int16:
            a = a.astype(np.float32)
        t = torch.from_numpy(a).to(dev)
    rho = engine.hu_to_density(t, knots)
    return rho if on_dev else rho.cpu().numpy()
