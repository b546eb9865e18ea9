"""Utilities either side of the dose path (reference core/utils.py), rebuilt without nibabel:

  * calculate_dvh (utils.py:233-262) on the GPU - min/max reduction + numpy-identical uniform-bin histogram,
    so a dose map that already lives in HBM is reduced to 2 x `bins` numbers instead of being copied out;
  * load_kernel / save_kernel: the binary `.dat` dose-kernel format (utils.py:17-51);
  * save_dose_map / load_dose_map: NIfTI-1 with the metadata JSON in header extension 44 (utils.py:53-152);
  * interpolate_timepoints (utils.py:154-191) on the GPU: scipy's interp1d is a linear map of the sampled volumes, so the
    host derives the J x T weights from the time points and ONE pass (pvd_weighted_combine) reads every sampled volume
    once and writes every interpolated volume once.
The file formats are host code (pyvoxeldosimetry_b200/io/), re-exported here under the reference's names.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

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


# ------------------------------------------------------------------------------------------------
# interpolate_timepoints (reference core/utils.py:154-191)

def _bspline_basis(t: np.ndarray, k: int, x: float, mu: int) -> np.ndarray:
    """The k+1 B-splines of degree k that are non-zero on the knot interval [t[mu], t[mu+1]) evaluated at x (de Boor /
    Cox recurrence).  x may lie outside the interval: the result is then the polynomial extension of that piece."""
    b = np.zeros(k + 1)
    b[0] = 1.0
    left, right = np.zeros(k + 1), np.zeros(k + 1)
    for j in range(1, k + 1):
        left[j], right[j] = x - t[mu + 1 - j], t[mu + j] - x
        saved = 0.0
        for r in range(j):
            term = b[r] / (right[r + 1] + left[j - r])
            b[r] = saved + right[r + 1] * term
            saved = left[j - r] * term
        b[j] = saved
    return b


def _cubic_weights(x: np.ndarray, xn: np.ndarray) -> np.ndarray:
    """Weights of the not-a-knot cubic spline interpolant (what interp1d(kind='cubic') builds through make_interp_spline):
    B-spline basis on the knots [x0]*4 + x[2:-2] + [x_last]*4, collocation at the samples, evaluation at xn with the end
    pieces extended beyond the range."""
    n, k = len(x), 3
    t = np.concatenate(([x[0]] * (k + 1), x[2:-2], [x[-1]] * (k + 1)))

    def row(v: float) -> np.ndarray:
        mu = int(np.clip(np.searchsorted(t, v, side="right") - 1, k, n - 1))
        r = np.zeros(n)
        r[mu - k:mu + 1] = _bspline_basis(t, k, v, mu)
        return r

    colloc = np.array([row(v) for v in x])
    return np.array([row(v) for v in xn]) @ np.linalg.inv(colloc)


def interpolation_weights(time_points: Sequence[float], new_times: Sequence[float], method: str = "linear") -> np.ndarray:
    """W[j, t] with interpolated_j = sum_t W[j, t] * values[t] for scipy.interpolate.interp1d(time_points, values, axis=0,
    kind=method, bounds_error=False, fill_value='extrapolate')(new_times) - the call of the reference (utils.py:179-187).
    Columns refer to the samples in the caller's order (interp1d sorts by time itself).  float64.
      'linear'   two non-zero weights per row, the end segments extrapolate;
      'previous' one weight 1 (the last sample at or before the new time); a row of NaN before the first sample, the
                 last sample beyond the end - interp1d's behaviour for this kind;
      'next'     mirror image;  'nearest' the closer sample (ties to the earlier one);
      'cubic'    not-a-knot cubic spline (needs >= 4 samples), end polynomials beyond the range."""
    x = np.asarray(time_points, dtype=np.float64)
    xn = np.atleast_1d(np.asarray(new_times, dtype=np.float64))
    if x.ndim != 1:
        raise ValueError("the x array must have exactly one dimension.")
    order = np.argsort(x, kind="mergesort")
    xs = x[order]
    n, J = len(xs), len(xn)
    Ws = np.zeros((J, n))
    rows = np.arange(J)
    if method in ("linear", "slinear"):
        if n < 2:
            raise ValueError("x and y arrays must have at least 2 entries")
        hi = np.searchsorted(xs, xn).clip(1, n - 1).astype(int)
        lo = hi - 1
        with np.errstate(divide="ignore", invalid="ignore"):
            Ws[rows, hi] = (xn - xs[lo]) / (xs[hi] - xs[lo])
            Ws[rows, lo] = (xs[hi] - xn) / (xs[hi] - xs[lo])
    elif method == "previous":
        idx = np.searchsorted(np.nextafter(xs, -np.inf), xn, side="left").clip(1, n).astype(int)
        Ws[rows, idx - 1] = 1.0
        Ws[xn > xs[-1]] = 0.0
        Ws[xn > xs[-1], n - 1] = 1.0
        Ws[xn < xs[0]] = np.nan
    elif method == "next":
        idx = np.searchsorted(np.nextafter(xs, np.inf), xn, side="right").clip(0, n - 1).astype(int)
        Ws[rows, idx] = 1.0
        Ws[xn < xs[0]] = 0.0
        Ws[xn < xs[0], 0] = 1.0
        Ws[xn > xs[-1]] = np.nan
    elif method == "nearest":
        bds = xs / 2.0
        idx = np.searchsorted(bds[1:] + bds[:-1], xn, side="left").clip(0, n - 1).astype(int)
        Ws[rows, idx] = 1.0
    elif method == "cubic":
        if n < 4:
            raise ValueError("x and y arrays must have at least 4 entries")
        Ws = _cubic_weights(xs, xn)
    else:
        raise NotImplementedError(f"{method} is unsupported: use 'linear', 'cubic', 'previous', 'next' or 'nearest'.")
    W = np.zeros_like(Ws)
    W[:, order] = Ws
    return W


def interpolate_timepoints(time_points: List[float], values: List, new_times: List[float], method: str = "linear",
                           device=None) -> List:
    """Interpolate 3-D arrays across time points (reference core/utils.py:154-191, same signature and errors).
    values: ndarrays (-> float32 ndarrays come back) or CUDA tensors (-> CUDA tensors, nothing leaves the device)."""
    if len(time_points) != len(values):
        raise ValueError("Number of time points must match number of values")
    W = interpolation_weights(time_points, new_times, method)
    on_dev = isinstance(values[0], torch.Tensor) and values[0].is_cuda
    dev = engine.require_cuda(device if device is not None else (values[0].device if on_dev else None))
    vols = [engine.to_device_f32(v, dev) for v in values]
    outs = engine.weighted_combine(vols, W.tolist())
    return outs if on_dev else [o.cpu().numpy() for o in outs]
