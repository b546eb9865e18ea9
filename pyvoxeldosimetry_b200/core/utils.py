"""Utilities either side of the dose path (reference core/utils.py), rebuilt without nibabel:

  * calculate_dvh (utils.py:233-262) on the GPU - min/max reduction + numpy-identical uniform-bin histogram,
    so a dose map that already lives in HBM is reduced to 2 x `bins` numbers instead of being copied out;
  * load_kernel / save_kernel: the binary `.dat` dose-kernel format (utils.py:17-51);
  * save_dose_map / load_dose_map: NIfTI-1 with the metadata JSON in header extension 44 (utils.py:53-152).
The file formats are host code (pyvoxeldosimetry_b200/io/), re-exported here under the reference's names.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .. import engine
from ..io.kernel_dat import load_kernel, save_kernel  # noqa: F401
from ..io.nifti import load_dose_map, save_dose_map  # noqa: F401


def calculate_dvh(dose_map, roi_mask, bins: int = 1000, device=None) -> Tuple[np.ndarray, np.ndarray]:
    """Cumulative dose-volume histogram: (dose_bins = edges[1:], volume_fraction = 1 - cumsum(hist)/N).
    Same errors as the reference (shape mismatch, empty ROI).  Doses are binned in float32 (the dose path's
    output type) with numpy.histogram's own rule, so counts equal np.histogram(dose32[mask > 0], bins)."""
    if tuple(dose_map.shape) != tuple(roi_mask.shape):
        raise ValueError("Dose map and ROI mask must have same dimensions")
    bins = int(bins)
    if bins < 1:
        raise ValueError("`bins` must be positive")
    dev = engine.require_cuda(device if device is not None else (dose_map.device if isinstance(dose_map, torch.Tensor) and dose_map.is_cuda else None))
    dose = engine.to_device_f32(dose_map, dev)
    if isinstance(roi_mask, torch.Tensor):
        m = roi_mask.to(dev)
        mask = m.contiguous() if m.dtype in (torch.uint8, torch.float32) else (m > 0).to(torch.uint8).contiguous()
    else:
        mask = torch.from_numpy(np.ascontiguousarray(np.asarray(roi_mask) > 0).view(np.uint8)).to(dev)
    mn, mx, count = engine.roi_minmax(dose, mask)
    if count == 0:
        raise ValueError("ROI mask is empty")
    # the edge array exactly as np.histogram builds it for float32 data spanning [mn, mx]
    edges = np.histogram_bin_edges(np.array([mn, mx], dtype=np.float32), bins=bins)
    hist = engine.dvh_histogram(dose, mask, torch.from_numpy(np.ascontiguousarray(edges, dtype=np.float32)).to(dev)).cpu().numpy()
    cum_dvh = 1.0 - np.cumsum(hist) / count
    return edges[1:], cum_dvh
