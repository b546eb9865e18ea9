"""ActivitySampler - mirrors core/activity_sampler.py:8-79 and supplies the missing
``integrate_dose_rates`` the reference front door calls (core/dose_calculator.py:138).

The trapezoid accumulations run on the GPU (pvd_weighted_sum) as one fused weighted sum
out = sum_i w_i * map_i instead of T-1 full-volume Python passes.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from .. import engine
from .kernel_convolution import trapezoid_weights

UNIT_TO_SECONDS = {"hours": 3600.0, "minutes": 60.0, "seconds": 1.0}


class ActivitySampler:
    def __init__(self, half_life: float, units: str = "hours", device=None):
        self.half_life = half_life
        self.units = units
        self._validate_inputs()
        self._device = device

    def _validate_inputs(self) -> None:
        if self.half_life <= 0:
            raise ValueError("Half-life must be positive")
        if self.units not in UNIT_TO_SECONDS:
            raise ValueError("Invalid time units")

    def sample_timepoints(self, start_time: float, end_time: float, n_points: int = 10) -> np.ndarray:
        return np.linspace(start_time, end_time, n_points)

    # ------------------------------------------------------------------
    def _weighted(self, maps, weights):
        if len(maps) == 0:
            raise ValueError("No activity maps provided")
        dev = engine.require_cuda(self._device)
        on_dev = isinstance(maps[0], torch.Tensor) and maps[0].is_cuda
        vols = [engine.to_device_f32(m, dev) for m in maps]
        out = engine.weighted_sum(vols, weights)
        return out if on_dev else out.cpu().numpy()

    def integrate_activity(self, activity_maps: List, time_points: List[float], method: str = "trapezoid"):
        """A3: trapezoid in the caller's time unit - no unit conversion (activity_sampler.py:69-79)."""
        if method != "trapezoid":
            raise ValueError(f"Unknown integration method: {method}")
        if len(activity_maps) != len(time_points):
            raise ValueError("Number of activity maps must match number of time points")
        return self._weighted(activity_maps, trapezoid_weights(time_points, 1.0))

    def dose_rate_weights(self, time_points: List[float], integration_limit: Optional[float] = None) -> List[float]:
        """Trapezoid weights in seconds (the A2 convention, kernel_convolution.py:102) plus, when
        ``integration_limit`` (same unit as time_points) lies beyond the last point, a physical-decay
        tail of the last rate: int_{t_last}^{limit} exp(-lambda (t - t_last)) dt."""
        f = UNIT_TO_SECONDS[self.units]
        w = trapezoid_weights(time_points, f)
        if integration_limit is not None and len(time_points) and integration_limit > time_points[-1]:
            lam = math.log(2.0) / self.half_life
            w[-1] += (1.0 - math.exp(-lam * (integration_limit - time_points[-1]))) / lam * f
        return w

    def integrate_dose_rates(self, dose_rates: List, time_points: List[float], integration_limit: Optional[float] = None):
        if len(dose_rates) != len(time_points):
            raise ValueError("Number of dose-rate maps must match number of time points")
        return self._weighted(dose_rates, self.dose_rate_weights(time_points, integration_limit))
