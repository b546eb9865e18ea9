"""Time-integration accumulation on the GPU: the closed-form integral of a fitted mono-exponential
(TimeCurveFitting._calculate_accumulated_dose, reference time_integration/curve_fitting.py:74-84).

Only the accumulation is in scope; the per-voxel scipy curve_fit loop
(curve_fitting.py:46-59) is the "next" row of section 8f and is not rebuilt here.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import engine


class TimeCurveFitting:
    def __init__(self, half_life: float, device=None):
        self.half_life = half_life
        self.decay_constant = float(np.log(2) / half_life)  # curve_fitting.py:17
        self._device = device

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

    def fit_time_activity_curve(self, *a, **k):
        raise NotImplementedError("per-voxel curve fitting is outside the rebuilt hot path (SURVEY.md section 8f rank 1)")
