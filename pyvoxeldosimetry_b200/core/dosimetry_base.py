"""Calculator plugin interface - mirrors DosimetryCalculator (reference core/dosimetry_base.py:8-71).

Same constructor arguments, abstract methods and error behaviour (TypeError for a non-string
radionuclide, dosimetry_base.py:28-31) so calculators written against the reference ABC plug in here.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Tuple

import numpy as np


class DosimetryCalculator(ABC):
    def __init__(self, radionuclide: str, tissue_composition: Any, config: Optional[Dict[str, Any]] = None):
        self.radionuclide = radionuclide
        self.tissue_composition = tissue_composition
        self.config = dict(config) if config else {}
        self._validate_inputs()

    def _validate_inputs(self) -> None:
        if not isinstance(self.radionuclide, str):
            raise TypeError("Radionuclide must be a string")

    @abstractmethod
    def calculate_dose_rate(self, activity_map: np.ndarray, voxel_size: Tuple[float, float, float]) -> np.ndarray:
        """3-D activity (Bq) -> 3-D dose rate (Gy/s)."""

    @abstractmethod
    def calculate_absorbed_dose(self, activity_maps: List[np.ndarray], time_points: List[float],
                                voxel_size: Tuple[float, float, float]) -> np.ndarray:
        """Time series of activity maps (time points in hours) -> absorbed dose (Gy)."""

    def get_config(self) -> Dict[str, Any]:
        return self.config.copy()
