"""Nuclide and tissue constants used by the dose-voxel-kernel generators.

Values restate the reference's data files (citations per entry); they are data, not code:
  Y90   data/dose_kernels/Y90/Y90.json:16,22-24      Lu177 data/dose_kernels/Lu177/Lu177.json:16,21-26
  Ga68  data/dose_kernels/Ga68/Ga68.json:16,22-27   F18 data/dose_kernels/F18/F18.json:16,22-24
Tissue tables: data/dose_kernels/y90_kernel.py:59-90, lu177_kernel.py:90-126, ga68_kernel.py:85-106.
The reference's Lu177/Ga68 JSON files carry `//` comments and cannot be parsed by json.load
(SURVEY.md Appendix B2), which is why the constants live in Python here.
"""
from __future__ import annotations

NUCLIDES = {
    "Y90": {"name": "Yttrium-90", "symbol": "Y90", "half_life": 64.1, "beta_max": 2.280, "beta_mean": 0.934,
            "bremsstrahlung": True, "default_grid": (201, 201, 201)},
    "Lu177": {"name": "Lutetium-177", "symbol": "Lu177", "half_life": 161.52, "beta_max": (0.498, 0.385, 0.177),
              "beta_abundance": (0.795, 0.089, 0.116), "gamma_lines": ((0.208, 0.111), (0.113, 0.062)),
              "default_grid": (81, 81, 81)},
    "F18": {"name": "Fluorine-18", "symbol": "F18", "half_life": 1.8295, "beta_max": 0.634, "annihilation": 0.511,
            "gamma_lines": (), "default_grid": (101, 101, 101)},
    "Ga68": {"name": "Gallium-68", "symbol": "Ga68", "half_life": 1.128, "beta_max": 1.899, "annihilation": 0.511,
             "gamma_lines": ((1.077, 0.03),), "default_grid": (151, 151, 151)},
}

# nuclides the reference registers (kernel_factory.py:15-22) but whose generators cannot run
# (SURVEY.md Appendix B3): Tb161 is abstract, Ac225 calls an undefined method.  (F18's helpers are `pass`
# in the reference, f18_kernel.py:38-46; it is supplied here with the positron form of the Ga68 generator.)
REFERENCE_BROKEN = ("Tb161", "Ac225")

TISSUES = {
    "water": {"density": 1.0, "effective_Z": 7.42, "stopping_power_ratio": 1.0, "mu_by_rho": 0.096},
    "lung": {"density": 0.26, "effective_Z": 7.41, "stopping_power_ratio": 1.04, "mu_by_rho": 0.095},
    "soft_tissue": {"density": 1.04, "effective_Z": 7.46, "stopping_power_ratio": 1.04, "mu_by_rho": 0.097},
    "bone": {"density": 1.85, "effective_Z": 13.8, "stopping_power_ratio": 1.15, "mu_by_rho": 0.110},
    "iodine_contrast": {"density": 1.30, "effective_Z": 53.0, "stopping_power_ratio": 1.12, "mu_by_rho": 0.245},
}
Y90_TISSUE_SCALE = {"water": 1.0, "bone": 1.15, "lung": 1.04, "soft_tissue": 1.04, "iodine_contrast": 1.12}
GA68_TISSUE_FACTOR = {"water": 1.0, "lung": 0.3, "soft_tissue": 1.04, "bone": 1.6, "iodine_contrast": 1.3}
GA68_MU_511 = {"water": 0.096, "lung": 0.029, "soft_tissue": 0.099, "bone": 0.172, "iodine_contrast": 0.158}


def tissue_props(tissue: str) -> dict:
    """Unknown tissue falls back to water, like the reference (y90_kernel.py:91, lu177_kernel.py:127)."""
    return TISSUES.get(tissue, TISSUES["water"])
