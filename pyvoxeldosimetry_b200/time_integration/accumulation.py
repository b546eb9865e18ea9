"""Time integration on the GPU: per-voxel mono-exponential fit of the time-activity curve and its closed-form
integral (reference time_integration/curve_fitting.py: TimeCurveFitting.fit_time_activity_curve :19-65,
_calculate_accumulated_dose :74-84).

The reference fits one voxel at a time with scipy.optimize.curve_fit (a Python loop: hours at 256^3); here the
same weighted least-squares problem is solved for every voxel in one kernel (pvd_monoexp_fit: damped
Gauss-Newton from the reference's own start point, then Newton on the variable-projection condition) and the
integral is fused into the same pass.  The result is the minimiser curve_fit converges to (to ~1e-6 for
well-posed curves); voxels whose fit is not finite get [0, decay_constant] like the reference's except-branch.
The "advanced" multi-model fits of the reference (fit_time_activity_curve_advanced) are outside the hot path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import engine


class TimeCurveFitting:
    def __init__(self, half_life: float, device=None):
        self.half_life = half_life
        self.decay_constant = float(np.log(2) / half_life)  # curve_fitting.py:17
        self._device = device

    def fit_time_activity_curve(self, times: Sequence[float], activities: List, weight_factors: Optional[Sequence[float]] = None):
        """-> (fitted_params [2, *shape] = (A0, lambda), accumulated [*shape]) as the reference returns
        (curve_fitting.py:64-65).  Host arrays in -> host float32 arrays out; CUDA tensors in -> CUDA tensors out."""
        times = [float(t) for t in np.asarray(times, dtype=np.float64).ravel()]
        if len(activities) != len(times):
            raise ValueError("Number of activity maps must match number of time points")
        if len(times) < 2:
            raise ValueError("a mono-exponential fit needs at least two time points")
        if weight_factors is not None and len(weight_factors) != len(times):
            raise ValueError("Number of weight factors must match number of time points")
        dev = engine.require_cuda(self._device)
        on_dev = isinstance(activities[0], torch.Tensor) and activities[0].is_cuda
        vols = [engine.to_device_f32(a, dev) for a in activities]
        shape = tuple(vols[0].shape)
        if any(tuple(v.shape) != shape for v in vols):
            raise ValueError("All activity maps must have the same shape")
        w = None if weight_factors is None else [float(x) for x in weight_factors]
        params, acc = engine.monoexp_fit(vols, times, w, self.decay_constant, 100.0 * self.half_life)
        if on_dev:
            return params, acc
        return params.cpu().numpy(), acc.cpu().numpy()

    def _calculate_accumulated_dose(self, fitted_params, integration_limit: Optional[float] = None):
        """fitted_params[0] = A0, fitted_params[1] = lambda (per voxel).  Returns A0/lambda*(1-exp(-lambda*T)),
        T defaulting to 100 half-lives (curve_fitting.py:78-79).  Host in -> host out, CUDA in -> CUDA out."""
        if integration_limit is None:
            integration_limit = 100 * self.half_life
        dev = engine.require_cuda(self._device)
        on_dev = isinstance(fitted_params, torch.Tensor) and fitted_params.is_cuda
        p = engine.to_device_f32(fitted_params, dev)
        out = engine.monoexp_integral(p[0].contiguous(), p[1].contiguous(), float(integration_limit))
        return out if on_dev else out.cpu().numpy()
