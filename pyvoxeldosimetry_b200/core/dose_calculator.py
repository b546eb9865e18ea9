"""DoseCalculator front door - same constructor / calculate_dose signature and routing as the reference
(core/dose_calculator.py:32-202), restricted to the kernel-convolution method (the hot path this
package rebuilds).  Dead branches of the reference are repaired:
  * `half_life` defaults to the nuclide's physical half-life instead of 0.0 -> ValueError;
  * the two methods the reference calls but never defines exist here;
  * `tissue_densities` is honoured (voxel-wise density correction fused into the last FFT pass);
  * a single time point fills `absorbed_dose` with the physical-decay integral of the dose rate
    (so examples/single_timepoint_y90_physical_decay.py runs) unless strict_reference is set.
Time units (recorded in result.metadata['time_unit_of_integration']): 'activity' mode integrates in the CALLER's unit
(hours by default: activity_sampler.py:74-78 applies no conversion, so do we), 'dose_rate' mode and the single-timepoint
physical-decay dose integrate in SECONDS (the x3600 of kernel_convolution.py:102) - the same inputs differ by the
unit factor between the modes, as they would in the reference had its missing methods existed.
In 'activity' mode result.dose_rate_maps is filled (T extra convolutions) because the reference's own example indexes
it (examples/time_integrated_dose.py:110); config['return_dose_rate_maps'] = False restores the reference's [].
Image registration (SimpleITK, core/image_registration.py) is out of scope: maps are taken as aligned
unless the caller supplies config['registration'] = callable(fixed, moving, spacing) -> aligned.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from ..data.nuclides import NUCLIDES
from .activity_sampler import ActivitySampler
from .kernel_convolution import KernelConvolutionCalculator, trapezoid_weights


class DoseCalculationResult:
    """Same four fields, positional order and repr as the reference's dataclass (core/dose_calculator.py:24-30).
    `absorbed_dose` may be handed over as a zero-argument callable: it is then evaluated on first access - the
    single-timepoint physical-decay dose is a full-volume product the caller often never reads (the reference itself
    returns None there), and a 105 M voxel NumPy multiply would otherwise dominate the call."""

    __slots__ = ("_absorbed", "dose_rate_maps", "time_points", "metadata")

    def __init__(self, absorbed_dose, dose_rate_maps: List[np.ndarray], time_points: List[float], metadata: Dict[str, Any]):
        self._absorbed = absorbed_dose
        self.dose_rate_maps = dose_rate_maps
        self.time_points = time_points
        self.metadata = metadata

    @property
    def absorbed_dose(self) -> Optional[np.ndarray]:
        if callable(self._absorbed):
            self._absorbed = self._absorbed()
        return self._absorbed

    @absorbed_dose.setter
    def absorbed_dose(self, value) -> None:
        self._absorbed = value

    def __repr__(self) -> str:
        return (f"DoseCalculationResult(absorbed_dose={self.absorbed_dose!r}, dose_rate_maps={self.dose_rate_maps!r}, "
                f"time_points={self.time_points!r}, metadata={self.metadata!r})")

    def __eq__(self, other) -> bool:
        if not isinstance(other, DoseCalculationResult):
            return NotImplemented
        return (np.array_equal(self.absorbed_dose, other.absorbed_dose) and len(self.dose_rate_maps) == len(other.dose_rate_maps)
                and all(np.array_equal(a, b) for a, b in zip(self.dose_rate_maps, other.dose_rate_maps))
                and list(self.time_points) == list(other.time_points) and self.metadata == other.metadata)


class DoseCalculator:
    def __init__(self, radionuclide: str, method: str = "kernel", config: Optional[Dict[str, Any]] = None):
        self.radionuclide = radionuclide
        self.method = method.lower()
        self.config = dict(config) if config else {}
        if self.method == "kernel":
            self.calculator = KernelConvolutionCalculator(
                radionuclide=radionuclide,
                tissue_name=self.config.get("tissue_name", "water"),
                kernel_resolution=self.config.get("kernel_resolution", 1.0),
                config=self.config,
            )
        elif self.method in ("gpu_monte_carlo", "gate_monte_carlo"):
            raise NotImplementedError(
                f"method {method!r} is outside this package: only the kernel-convolution dose path is rebuilt for B200")
        else:
            raise ValueError(f"Unsupported calculation method: {method}")
        half_life = self.config.get("half_life")
        if half_life is None:
            half_life = NUCLIDES.get(radionuclide, {}).get("half_life", 0.0)
        self.activity_sampler = ActivitySampler(half_life=half_life, units=self.config.get("time_units", "hours"),
                                                device=self.calculator.device)
        self.image_registration = self.config.get("registration")  # optional callable

    # ------------------------------------------------------------------
    def calculate_dose(self, activity_maps: Optional[List[np.ndarray]] = None, time_points: Optional[List[float]] = None,
                       voxel_size: Tuple[float, float, float] = (1.0, 1.0, 1.0), tissue_densities: Optional[np.ndarray] = None,
                       accumulated_activity: Optional[np.ndarray] = None, integration_mode: str = "activity",
                       integration_limit: Optional[float] = None) -> DoseCalculationResult:
        if len(voxel_size) != 3:
            raise ValueError("voxel_size must be a tuple of length 3.")
        calc = self.calculator

        if accumulated_activity is not None:
            dose = calc.calculate_absorbed_dose_from_accumulated(accumulated_activity, voxel_size, tissue_densities)
            return DoseCalculationResult(dose, [], [], {"mode": "accumulated_activity"})

        if activity_maps is not None and time_points is not None and len(activity_maps) > 1:
            self._validate_inputs(activity_maps, time_points, voxel_size)
            maps = self._align_activity_maps(activity_maps, voxel_size)
            if integration_mode == "activity":
                # integrate activity (caller's time unit, activity_sampler.py:74-78), then one convolution
                w = trapezoid_weights(time_points, 1.0)
                dose = calc.calculate_weighted(maps, w, voxel_size, tissue_densities)
                rates = []
                if not calc.strict_reference and self.config.get("return_dose_rate_maps", True):
                    # the reference returns [] here yet its own example plots result.dose_rate_maps[i]
                    # (examples/time_integrated_dose.py:110); on the GPU the T extra convolutions are cheap
                    rates = [calc.calculate_dose_rate(a, voxel_size, tissue_densities) for a in maps]
                return DoseCalculationResult(dose, rates, time_points, {"mode": "multi_timepoint_activity",
                                                                        "time_unit_of_integration": self.activity_sampler.units})
            if integration_mode == "dose_rate":
                rates = [calc.calculate_dose_rate(a, voxel_size, tissue_densities) for a in maps]
                dose = self.activity_sampler.integrate_dose_rates(rates, time_points, integration_limit)
                return DoseCalculationResult(dose, rates, time_points, {"mode": "multi_timepoint_doserate",
                                                                        "time_unit_of_integration": "seconds"})
            raise ValueError(f"Unknown integration_mode: {integration_mode}")

        if activity_maps is not None and len(activity_maps) == 1:
            self._validate_inputs(activity_maps, time_points or [0], voxel_size)
            rate = calc.calculate_dose_rate(activity_maps[0], voxel_size, tissue_densities)
            absorbed = None
            if not calc.strict_reference:
                # physical-decay integral of the dose rate from the scan time on: rate * T_half / ln 2
                f = {"hours": 3600.0, "minutes": 60.0, "seconds": 1.0}[self.activity_sampler.units]
                factor = self.activity_sampler.half_life * f / math.log(2.0)
                if self.config.get("decay_correct_to_t0") and time_points:
                    factor *= math.exp(math.log(2.0) * float(time_points[0]) / self.activity_sampler.half_life)
                absorbed = (lambda: rate * np.float32(factor)) if rate.dtype == np.float32 else (lambda: rate * factor)
            return DoseCalculationResult(absorbed, [rate], time_points or [], {"mode": "single_timepoint",
                                                                             "time_unit_of_integration": "seconds"})

        raise ValueError("Invalid input for dose calculation.")

    # ------------------------------------------------------------------
    def _validate_inputs(self, activity_maps, time_points, voxel_size) -> None:
        if not activity_maps:
            raise ValueError("No activity maps provided")
        if len(activity_maps) != len(time_points):
            raise ValueError("Number of activity maps must match number of time points")
        if len(voxel_size) != 3:
            raise ValueError("Voxel size must be 3D")
        shape = activity_maps[0].shape
        if not all(m.shape == shape for m in activity_maps):
            raise ValueError("All activity maps must have the same dimensions")

    def _align_activity_maps(self, activity_maps, voxel_size):
        if len(activity_maps) == 1 or self.image_registration is None:
            return list(activity_maps)
        ref = activity_maps[0]
        return [ref] + [self.image_registration(ref, m, voxel_size) for m in activity_maps[1:]]
