"""Dose-voxel-kernel generators: same classes / methods as the reference
(BaseKernelGenerator data/dose_kernels/base_kernel.py:22-84, Y90KernelGenerator y90_kernel.py:7-162,
Lu177KernelGenerator lu177_kernel.py:38-190, Ga68KernelGenerator ga68_kernel.py:7-106), but the
voxel-grid evaluation runs on the GPU (pvd_kernel_eval_radial): the Python side only turns nuclide /
tissue constants into the terms of the radial model.

Differences from the reference, all deliberate:
  * the r = 0 voxel is finite (photon term := 0 there); the reference's Y90/Ga68 kernels hold a NaN
    (y90_kernel.py:134-138) that turns every dose map into NaN;
  * ``voxel_size`` may be a 3-tuple (anisotropic image grids, A10); a scalar reproduces the reference;
  * no PNG / JSON side effects when a kernel is generated (base_kernel.py:43-84).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np

from ..nuclides import GA68_MU_511, GA68_TISSUE_FACTOR, NUCLIDES, Y90_TISSUE_SCALE, tissue_props

Spacing = Union[float, Sequence[float]]


def _spacing3(voxel_size: Spacing) -> Tuple[float, float, float]:
    if np.isscalar(voxel_size):
        v = float(voxel_size)
        sp = (v, v, v)
    else:
        sp = tuple(float(v) for v in voxel_size)
        if len(sp) != 3:
            raise ValueError("voxel_size must be a scalar or have length 3")
    if not all(v > 0 for v in sp):
        raise ValueError("voxel_size must be positive")
    return sp  # type: ignore[return-value]


class BaseKernelGenerator(ABC):
    nuclide: str = ""

    def __init__(self, tissue_type: str):
        self.config = {"nuclide": dict(NUCLIDES[self.nuclide])}
        self.tissue_type = tissue_type
        self.tissue_properties = tissue_props(tissue_type)

    @abstractmethod
    def radial_terms(self) -> Tuple[List[Tuple[float, float]], List[Tuple[float, float]], float]:
        """-> (beta terms [(range_mm, amplitude)], photon terms [(mu_per_cm, amplitude)], final scaling)."""

    def generate_kernel_device(self, voxel_size: Spacing, grid_size: Sequence[int], device=None):
        """Evaluate on the GPU, leave the float32 kernel on the device (torch.Tensor)."""
        from ... import engine

        beta, phot, scale = self.radial_terms()
        grid = tuple(int(g) for g in grid_size)
        if len(grid) != 3 or min(grid) < 1:
            raise ValueError("grid_size must be three positive integers")
        return engine.kernel_eval_radial(beta, phot, scale, _spacing3(voxel_size), grid, device)

    def generate_kernel(self, voxel_size: Spacing, grid_size: Sequence[int]) -> np.ndarray:
        """Reference signature (base_kernel.py:38-41): returns a host float64 ndarray."""
        return self.generate_kernel_device(voxel_size, grid_size).cpu().numpy().astype(np.float64)

    def save_kernel(self, kernel: np.ndarray, output_dir) -> None:
        """np.save only (the reference also renders a PNG and a JSON side-car, base_kernel.py:43-84)."""
        from pathlib import Path

        np.save(Path(output_dir) / f"{self.nuclide}_{self.tissue_type}_kernel.npy", kernel)


def _beta_term(range_coeff: float, energy: float, props: dict, amplitude: float = 1.0) -> Tuple[float, float]:
    # CSDA-like range scaled by 1/(rho*S); amplitude rho*S  (y90_kernel.py:99-115, lu177_kernel.py:137-152)
    rho, spr = props["density"], props["stopping_power_ratio"]
    rng = range_coeff * energy ** 1.5 * (1.0 / rho) * (1.0 / spr)
    return rng, amplitude * rho * spr


class Y90KernelGenerator(BaseKernelGenerator):
    nuclide = "Y90"

    def __init__(self, tissue_type: str):
        super().__init__(tissue_type)
        self.beta_max_energy = NUCLIDES["Y90"]["beta_max"]
        self.beta_mean_energy = NUCLIDES["Y90"]["beta_mean"]

    def radial_terms(self):
        p = self.tissue_properties
        beta = [_beta_term(11.0, self.beta_max_energy, p)]
        rel_yield = (p["effective_Z"] / 7.42) ** 2            # y90_kernel.py:128
        mu = 0.096 * (p["density"] / 1.0)                      # y90_kernel.py:142-146
        phot = [(mu, 0.015 * rel_yield * p["density"])]        # y90_kernel.py:134-138
        return beta, phot, Y90_TISSUE_SCALE.get(self.tissue_type, 1.0)


class Lu177KernelGenerator(BaseKernelGenerator):
    nuclide = "Lu177"

    def __init__(self, tissue_type: str):
        super().__init__(tissue_type)
        n = NUCLIDES["Lu177"]
        self.beta_energies, self.beta_abundances, self.gamma_lines = n["beta_max"], n["beta_abundance"], n["gamma_lines"]

    def radial_terms(self):
        p = self.tissue_properties
        beta = [_beta_term(5.0, e, p, ab) for e, ab in zip(self.beta_energies, self.beta_abundances)]
        phot = [(p["density"] * p["mu_by_rho"] * (0.2 / e) ** 3.2, inten) for e, inten in self.gamma_lines]
        return beta, phot, 1.0  # no final tissue factor: lu177_kernel.py:186-190 is never called


class Ga68KernelGenerator(BaseKernelGenerator):
    nuclide = "Ga68"

    def __init__(self, tissue_type: str):
        super().__init__(tissue_type)
        n = NUCLIDES["Ga68"]
        self.beta_max_energy, self.gamma_lines = n["beta_max"], n["gamma_lines"]

    def radial_terms(self):
        f = GA68_TISSUE_FACTOR.get(self.tissue_type, 1.0)      # ga68_kernel.py:85-94
        mu511 = GA68_MU_511.get(self.tissue_type, 0.096)       # ga68_kernel.py:96-106
        beta = [(9.0 * self.beta_max_energy ** 1.5 * f, 1.0)]  # ga68_kernel.py:59-70
        phot = [(mu511 * (0.511 / 0.511) ** 3.2, 1.0)]         # annihilation photons, ga68_kernel.py:72-75
        phot += [(mu511 * (0.511 / e) ** 3.2, inten) for e, inten in self.gamma_lines]
        return beta, phot, f


class F18KernelGenerator(Ga68KernelGenerator):
    """F-18: positron range + 511 keV annihilation photons.  The reference's generator cannot run
    (_calculate_beta_range / _dose_point_value are `pass`, f18_kernel.py:38-46; its docstring states the
    intent: "positron range and annihilation photon contributions"), so this uses the positron form of the
    Ga68 generator with E_max = 0.634 MeV (F18/F18.json:22).  PARITY UNPINNED - no reference values exist."""
    nuclide = "F18"

    def __init__(self, tissue_type: str):
        BaseKernelGenerator.__init__(self, tissue_type)
        n = NUCLIDES["F18"]
        self.beta_max_energy, self.gamma_lines = n["beta_max"], n["gamma_lines"]


GENERATORS: Dict[str, type] = {"Y90": Y90KernelGenerator, "Lu177": Lu177KernelGenerator, "Ga68": Ga68KernelGenerator,
                               "F18": F18KernelGenerator}
