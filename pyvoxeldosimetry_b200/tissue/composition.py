"""Tissue composition from CT on the GPU (reference tissue/composition.py:37-93).

`TissueComposition.calculate_composition(ct_image, handle_artifacts=True)` returns the same dict of 0/1 maps
(air, lung, soft_tissue, bone, water by HU range, :40-46,63-67) after the same metal-artifact handling
(voxels > 2000 HU replaced by a sigma=1 Gaussian smoothing of the image with those voxels zeroed, :73-93).
One kernel pass (pvd_ct_prepare) produces the corrected HU volume, the tissue bit labels and - for the density
correction of the dose path - the mass-density map.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch

from .. import engine
from .density import HU_KNOTS

METAL_THRESHOLD_HU = 2000.0  # composition.py:84


class TissueComposition:
    def __init__(self, config: Optional[Dict[str, Any]] = None, device=None):
        self.config = config or {}
        self.tissue_types = {  # composition.py:40-46 (insertion order = bit order of the device labels)
            "air": {"hu_range": (-1000, -900)},
            "lung": {"hu_range": (-900, -500)},
            "soft_tissue": {"hu_range": (-100, 100)},
            "bone": {"hu_range": (300, 3000)},
            "water": {"hu_range": (-10, 10)},
        }
        self._device = device

    def _run(self, ct_image, handle_artifacts: bool, knots=None):
        dev = engine.require_cuda(self._device)
        hu = engine.to_device_f32(ct_image, dev)
        ranges = [t["hu_range"] for t in self.tissue_types.values()]
        thr = METAL_THRESHOLD_HU if handle_artifacts else float("inf")
        return engine.ct_prepare(hu, thr, knots, ranges, want_corrected=True)

    def calculate_composition(self, ct_image, handle_artifacts: bool = True) -> Dict[str, Any]:
        """Dict of 0/1 maps per tissue class.  Host array in -> float64 host arrays (as the reference's
        mask.astype(float)); CUDA tensor in -> float32 CUDA tensors."""
        on_dev = isinstance(ct_image, torch.Tensor) and ct_image.is_cuda
        _, _, labels = self._run(ct_image, handle_artifacts)
        out = {}
        host = None if on_dev else labels.cpu().numpy()
        for bit, name in enumerate(self.tissue_types):
            if on_dev:
                out[name] = ((labels >> bit) & 1).to(torch.float32)
            else:
                out[name] = ((host >> bit) & 1).astype(float)
        return out

    def _handle_artifacts(self, ct_image):
        """Metal-artifact handled CT (composition.py:73-93).  Host in -> host out, CUDA in -> CUDA out."""
        on_dev = isinstance(ct_image, torch.Tensor) and ct_image.is_cuda
        corrected, _, _ = self._run(ct_image, True)
        return corrected if on_dev else corrected.cpu().numpy()

    def density_map(self, ct_image, handle_artifacts: bool = True, knots=HU_KNOTS):
        """Mass density (g/cm3) of the artifact-handled CT: the `tissue_densities` input of the dose path."""
        on_dev = isinstance(ct_image, torch.Tensor) and ct_image.is_cuda
        _, rho, _ = self._run(ct_image, handle_artifacts, knots)
        return rho if on_dev else rho.cpu().numpy()
