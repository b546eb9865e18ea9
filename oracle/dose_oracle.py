"""CPU oracle for the kernel-convolution dose path.  TEST INFRASTRUCTURE ONLY.

This file is a float64 NumPy restatement of the reference's algorithm for the hot path
 plus the float64 definition of the two spec-defined
steps that have no reference code (A9 density correction, A10 anisotropic kernel grid).
It is the *checker* for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
``pyvoxeldosimetry_b200`` never does.

Pinning status
--------------
* A1/A2/A3 (convolution + trapezoids): the arithmetic lives in ``numpy.fft`` (pocketfft, not
  vendored by the reference; the reference pins only ``numpy>=1.20.0`` in setup.py:8).  The
  restatement uses the identical NumPy calls, is cross-checked bit-for-bit against the real
  reference class imported under the stub recipe (``oracle/gen_golden.py``) and against a brute
  force O(N*K) definition.  Golden vectors produced by the real reference are committed under
  ``tests/golden/``.  The reference itself ships NO tests / golden vectors.
* A5/A6 (Y90 / Lu177 dose-voxel kernels): pinned against the real generators, every voxel
  except the r = 0 voxel of Y90 where the reference yields NaN (y90_kernel.py:134-138).
* A9 density correction, A10 anisotropic spacing, single-timepoint physical-decay dose:
  **parity unpinned** - no reference code exists; the formulas below are the specification.

* Section 8f rows (mono-exponential fit, CT artifact handling / tissue classes, DVH, `.dat` kernels): the
  restatements call the same SciPy / NumPy routines as the reference and are pinned against the real
  reference classes by ``oracle/gen_golden.py`` (``tests/golden/next_ref.npz``).  The NIfTI writer of the
  reference needs nibabel, which is absent here: **NIfTI parity unpinned** (checked against the NIfTI-1
  field layout and by round trip only).

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# A1  KernelConvolutionCalculator.calculate_dose_rate   core/kernel_convolution.py:71-74
# --------------------------------------------------------------------------------------


def conv_reference(activity: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """Literal restatement of core/kernel_convolution.py:71-74.

    Circular convolution over ``activity.shape`` with the kernel cropped / zero padded *at the
    origin* (``np.fft.fftn(kernel, s=shape)``), complex128, ``.real`` taken.
    """
    activity = np.asarray(activity, dtype=np.float64)
    kernel = np.asarray(kernel, dtype=np.float64)
    return np.fft.ifftn(np.fft.fftn(activity) * np.fft.fftn(kernel, activity.shape)).real


def conv_reference_fast(activity: np.ndarray, kernel: np.ndarray, workers: int = -1) -> np.ndarray:
    """Same mathematics as :func:`conv_reference` through scipy.fft real transforms with all
    host threads (SURVEY.md section 8d "Baseline B").  Used only as a fast checker for the
    big shapes; agrees with the literal form to ~1e-15 of peak."""
    import scipy.fft as sfft

    activity = np.asarray(activity, dtype=np.float64)
    kernel = np.asarray(kernel, dtype=np.float64)
    shape = activity.shape
    kt = kernel[tuple(slice(0, min(k, n)) for k, n in zip(kernel.shape, shape))]
    fa = sfft.rfftn(activity, workers=workers)
    fk = sfft.rfftn(kt, s=shape, workers=workers)
    fa *= fk
    return sfft.irfftn(fa, s=shape, workers=workers)


def crop_pad_kernel(kernel: np.ndarray, shape: Sequence[int]) -> np.ndarray:
    """k~ of SURVEY Appendix A.1: what ``np.fft.fftn(kernel, s=shape)`` transforms."""
    out = np.zeros(shape, dtype=np.float64)
    sl = tuple(slice(0, min(k, n)) for k, n in zip(kernel.shape, shape))
    out[sl] = np.asarray(kernel, dtype=np.float64)[sl]
    return out


def conv_bruteforce(activity: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """Definition of the reference operator without any FFT:
    d[n] = sum_j k~[j] * a[(n - j) mod N].  O(N*K) - small grids only."""
    a = np.asarray(activity, dtype=np.float64)
    kt = crop_pad_kernel(kernel, a.shape)
    out = np.zeros_like(a)
    nz = np.argwhere(kt != 0.0)
    for j0, j1, j2 in nz:
        out += kt[j0, j1, j2] * np.roll(a, shift=(j0, j1, j2), axis=(0, 1, 2))
    return out


def kernel_centre(kshape: Sequence[int]) -> Tuple[int, int, int]:
    """Centre convention of the generators: ``s // 2`` (y90_kernel.py:34)."""
    return tuple(int(s) // 2 for s in kshape)  # type: ignore[return-value]


def conv_same(activity: np.ndarray, kernel: np.ndarray, fast: bool = False) -> np.ndarray:
    """Zero-boundary, centred ("same") convolution expressed THROUGH the reference operator
    (SURVEY Appendix A.5): same = ref(pad0(a, N+K-1), k)[c : c+N], c = K // 2."""
    a = np.asarray(activity, dtype=np.float64)
    k = np.asarray(kernel, dtype=np.float64)
    big = tuple(n + kk - 1 for n, kk in zip(a.shape, k.shape))
    pad = np.zeros(big, dtype=np.float64)
    pad[: a.shape[0], : a.shape[1], : a.shape[2]] = a
    full = conv_reference_fast(pad, k) if fast else conv_reference(pad, k)
    c = kernel_centre(k.shape)
    return full[c[0] : c[0] + a.shape[0], c[1] : c[1] + a.shape[1], c[2] : c[2] + a.shape[2]]


# --------------------------------------------------------------------------------------
# A2  calculate_absorbed_dose   core/kernel_convolution.py:94-106
# A3  ActivitySampler._trapezoid_integration   core/activity_sampler.py:69-79
# --------------------------------------------------------------------------------------


def absorbed_dose_trapezoid(
    activity_maps: Sequence[np.ndarray], time_points: Sequence[float], kernel: np.ndarray, conv=conv_reference
) -> np.ndarray:
    """Literal loop of core/kernel_convolution.py:94-106 (hours -> seconds, x3600 at :102)."""
    maps = [np.asarray(m, dtype=np.float64) for m in activity_maps]
    dose_map = np.zeros_like(maps[0])
    dose_rates = [conv(m, kernel) for m in maps]
    for i in range(len(time_points) - 1):
        dt = (time_points[i + 1] - time_points[i]) * 3600
        avg = (dose_rates[i] + dose_rates[i + 1]) / 2
        dose_map += avg * dt
    return dose_map


def trapezoid_weights(time_points: Sequence[float], unit_factor: float = 1.0) -> np.ndarray:
    """Closed form of the two trapezoid loops: D = sum_i w_i f_i with
    w_0 = d_0/2, w_i = (d_{i-1}+d_i)/2, w_{T-1} = d_{T-2}/2, d_i = (t_{i+1}-t_i)*unit_factor.
    unit_factor = 3600 for A2 (kernel_convolution.py:102), 1 for A3 (activity_sampler.py:76)."""
    t = np.asarray(time_points, dtype=np.float64)
    w = np.zeros(len(t), dtype=np.float64)
    if len(t) >= 2:
        d = np.diff(t) * unit_factor
        w[:-1] += d / 2
        w[1:] += d / 2
    return w


def integrate_activity_trapezoid(activity_maps: Sequence[np.ndarray], time_points: Sequence[float]) -> np.ndarray:
    """Literal loop of core/activity_sampler.py:69-79 (no unit conversion)."""
    maps = [np.asarray(m, dtype=np.float64) for m in activity_maps]
    result = np.zeros_like(maps[0])
    for i in range(len(time_points) - 1):
        dt = time_points[i + 1] - time_points[i]
        result += (maps[i] + maps[i + 1]) / 2 * dt
    return result


def integrate_dose_rates(
    dose_rates: Sequence[np.ndarray],
    time_points: Sequence[float],
    integration_limit: Optional[float] = None,
    half_life: Optional[float] = None,
    unit_factor: float = 3600.0,
) -> np.ndarray:
    """Intended meaning of the missing ``ActivitySampler.integrate_dose_rates``
    (core/dose_calculator.py:138): A2's trapezoid on precomputed
    rates, plus (when ``integration_limit`` and ``half_life`` are given) a physical-decay tail of
    the last rate from t_last to ``integration_limit``.  Tail part is parity-unpinned."""
    rates = [np.asarray(r, dtype=np.float64) for r in dose_rates]
    w = trapezoid_weights(time_points, unit_factor)
    if integration_limit is not None and half_life is not None and integration_limit > time_points[-1]:
        lam = math.log(2.0) / half_life
        tail = (1.0 - math.exp(-lam * (integration_limit - time_points[-1]))) / lam
        w[-1] += tail * unit_factor
    out = np.zeros_like(rates[0])
    for wi, r in zip(w, rates):
        out += wi * r
    return out


# --------------------------------------------------------------------------------------
# A11 TimeCurveFitting._calculate_accumulated_dose  time_integration/curve_fitting.py:74-84
# --------------------------------------------------------------------------------------


def accumulated_activity_monoexp(
    A0: np.ndarray, lam: np.ndarray, half_life: float, integration_limit: Optional[float] = None
) -> np.ndarray:
    """A0/lambda * (1 - exp(-lambda * T_lim)), T_lim default 100*half_life (:78-79)."""
    if integration_limit is None:
        integration_limit = 100 * half_life
    A0 = np.asarray(A0, dtype=np.float64)
    lam = np.asarray(lam, dtype=np.float64)
    return A0 / lam * (1 - np.exp(-lam * integration_limit))


# --------------------------------------------------------------------------------------
# A5  Y90KernelGenerator   data/dose_kernels/y90_kernel.py:20-162, Y90/Y90.json:21-25
# A6  Lu177KernelGenerator data/dose_kernels/lu177_kernel.py:52-184, Lu177/Lu177.json:20-27
# --------------------------------------------------------------------------------------

# y90_kernel.py:59-90 (density, effective_Z, stopping_power_ratio) and lu177_kernel.py:90-126
# (same three + mu_by_rho).  Unknown tissue -> water (y90_kernel.py:91, lu177_kernel.py:127).
TISSUES = {
    "water": dict(density=1.0, effective_Z=7.42, stopping_power_ratio=1.0, mu_by_rho=0.096),
    "lung": dict(density=0.26, effective_Z=7.41, stopping_power_ratio=1.04, mu_by_rho=0.095),
    "soft_tissue": dict(density=1.04, effective_Z=7.46, stopping_power_ratio=1.04, mu_by_rho=0.097),
    "bone": dict(density=1.85, effective_Z=13.8, stopping_power_ratio=1.15, mu_by_rho=0.110),
    "iodine_contrast": dict(density=1.30, effective_Z=53.0, stopping_power_ratio=1.12, mu_by_rho=0.245),
}
# y90_kernel.py:148-162
Y90_TISSUE_SCALE = {"water": 1.0, "bone": 1.15, "lung": 1.04, "soft_tissue": 1.04, "iodine_contrast": 1.12}
Y90_BETA_MAX = 2.280  # Y90/Y90.json:22
Y90_HALF_LIFE_H = 64.1  # Y90/Y90.json:16
LU177_BETA_MAX = (0.498, 0.385, 0.177)  # Lu177/Lu177.json:21
LU177_BETA_AB = (0.795, 0.089, 0.116)  # Lu177/Lu177.json:22
LU177_GAMMAS = ((0.208, 0.111), (0.113, 0.062))  # (energy MeV, intensity) Lu177/Lu177.json:23-26
LU177_HALF_LIFE_H = 161.52  # Lu177/Lu177.json:16


def radial_grid(grid_size: Sequence[int], spacing: Sequence[float]) -> np.ndarray:
    """r = ||(idx - size//2) * spacing||  (y90_kernel.py:34-43 with per-axis spacing, A10)."""
    c = [s // 2 for s in grid_size]
    x, y, z = np.meshgrid(
        np.arange(grid_size[0]) - c[0], np.arange(grid_size[1]) - c[1], np.arange(grid_size[2]) - c[2], indexing="ij"
    )
    return np.sqrt((x * spacing[0]) ** 2 + (y * spacing[1]) ** 2 + (z * spacing[2]) ** 2)


def _beta_term(r: np.ndarray, range_coeff: float, energy: float, props: dict) -> np.ndarray:
    """y90_kernel.py:93-117 / lu177_kernel.py:129-154."""
    density_factor = props["density"]
    spr = props["stopping_power_ratio"]
    max_range = range_coeff * energy**1.5
    tissue_range = max_range * (1.0 / density_factor) * (1.0 / spr)
    out = np.zeros_like(r)
    mask = r <= tissue_range
    out[mask] = (1 - r[mask] / tissue_range) ** 2 * np.exp(-2 * r[mask] / tissue_range) * density_factor * spr
    return out


def y90_kernel(
    voxel_size, grid_size: Sequence[int], tissue: str = "water", centre: str = "finite"
) -> np.ndarray:
    """Restatement of Y90KernelGenerator.generate_kernel (y90_kernel.py:20-55).

    ``voxel_size`` scalar (reference) or 3 per-axis values (A10 extension).
    ``centre='finite'``: bremsstrahlung(0) := 0 (the masked form the same author uses in
    lu177_kernel.py:176-182); ``centre='reference'``: reproduce the NaN of y90_kernel.py:134-138.
    """
    sp = (voxel_size,) * 3 if np.isscalar(voxel_size) else tuple(voxel_size)
    props = TISSUES.get(tissue, TISSUES["water"])
    r = radial_grid(grid_size, sp)
    kernel = np.zeros(tuple(grid_size))
    kernel += _beta_term(r, 11.0, Y90_BETA_MAX, props)
    relative_yield = (props["effective_Z"] / 7.42) ** 2
    mu = 0.096 * (props["density"] / 1.0)
    if centre == "reference":
        with np.errstate(divide="ignore", invalid="ignore"):
            brems = 0.015 * relative_yield * props["density"] * np.exp(-mu * r / 10) / (4 * np.pi * r**2) * (r > 0)
    else:
        brems = np.zeros_like(r)
        m = r > 0
        brems[m] = 0.015 * relative_yield * props["density"] * np.exp(-mu * r[m] / 10) / (4 * np.pi * r[m] ** 2)
    kernel += brems
    kernel *= Y90_TISSUE_SCALE.get(tissue, 1.0)
    return kernel


def lu177_kernel(voxel_size, grid_size: Sequence[int], tissue: str = "water") -> np.ndarray:
    """Restatement of Lu177KernelGenerator.generate_kernel (lu177_kernel.py:52-86)."""
    sp = (voxel_size,) * 3 if np.isscalar(voxel_size) else tuple(voxel_size)
    props = TISSUES.get(tissue, TISSUES["water"])
    r = radial_grid(grid_size, sp)
    kernel = np.zeros(tuple(grid_size))
    for energy, abundance in zip(LU177_BETA_MAX, LU177_BETA_AB):
        kernel += abundance * _beta_term(r, 5.0, energy, props)
    for energy, intensity in LU177_GAMMAS:
        energy_factor = (0.2 / energy) ** 3.2
        mu = props["density"] * props["mu_by_rho"] * energy_factor
        g = np.zeros_like(r)
        m = r > 0
        g[m] = intensity * np.exp(-mu * r[m] / 10) / (4 * np.pi * (r[m] ** 2))
        kernel += g
    return kernel


def make_kernel(nuclide: str, voxel_size, grid_size, tissue="water") -> np.ndarray:
    if nuclide == "Y90":
        return y90_kernel(voxel_size, grid_size, tissue)
    if nuclide == "Lu177":
        return lu177_kernel(voxel_size, grid_size, tissue)
    raise ValueError(f"oracle has no generator for {nuclide}")


# --------------------------------------------------------------------------------------
# A9  density correction (NEW, spec-defined, parity unpinned)
# --------------------------------------------------------------------------------------

# Piecewise-linear HU -> mass density (g/cm3).  Knots follow the material table the reference
# ships for GATE (core/gate/data/HU_to_material.txt:5-20 with GateMaterials.db densities:
# Air 0.00129, Lung 0.26, Adipose 0.92, water 1.0, Muscle 1.05, SpineBone 1.42, RibBone 1.92).
HU_KNOTS = np.array(
    [
        (-1000.0, 0.00129),
        (-700.0, 0.26),
        (-100.0, 0.92),
        (0.0, 1.0),
        (40.0, 1.05),
        (350.0, 1.42),
        (1200.0, 1.92),
        (3000.0, 2.90),
    ],
    dtype=np.float64,
)


def hu_to_density(hu: np.ndarray, knots: np.ndarray = HU_KNOTS) -> np.ndarray:
    """Piecewise linear interpolation, clamped at both ends."""
    return np.interp(np.asarray(hu, dtype=np.float64), knots[:, 0], knots[:, 1])


def density_correct(
    dose: np.ndarray, rho: np.ndarray, rho_ref: float = 1.0, rho_min: float = 0.1, rho_cut: float = 0.0
) -> np.ndarray:
    """D_corr = D * rho_ref / max(rho, rho_min); voxels with rho < rho_cut are zeroed."""
    rho = np.asarray(rho, dtype=np.float64)
    out = np.asarray(dose, dtype=np.float64) * (rho_ref / np.maximum(rho, rho_min))
    if rho_cut > 0.0:
        out = np.where(rho < rho_cut, 0.0, out)
    return out


# --------------------------------------------------------------------------------------
# Slab decomposition reference - pure index logic, CPU checkable
# --------------------------------------------------------------------------------------


def slab_bounds(n0: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous split of axis 0 into ``world`` slabs (first ``n0 % world`` get one extra)."""
    base, rem = divmod(n0, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def conv_reference_slabbed(activity: np.ndarray, kernel: np.ndarray, world: int, conv=conv_reference) -> np.ndarray:
    """Reference-mode (circular) convolution computed slab by slab along axis 0 with a one-sided
    (K0e-1)-plane halo from the lower neighbour (circular) - overlap-save.  Must equal
    conv_reference(activity, kernel)."""
    a = np.asarray(activity, dtype=np.float64)
    n0 = a.shape[0]
    k = crop_pad_kernel(kernel, tuple(min(kk, n) for kk, n in zip(kernel.shape, a.shape)))
    h = k.shape[0] - 1
    out = np.empty_like(a)
    for lo, hi in slab_bounds(n0, world):
        idx = np.arange(lo - h, hi) % n0
        local = a[idx]
        # local circular conv along axis 0 of length (hi-lo)+h is exact for rows >= h;
        # axes 1,2 keep the global circular semantics because they are not split.
        d = conv(local, k)
        out[lo:hi] = d[h:]
    return out


def conv_same_slabbed(activity: np.ndarray, kernel: np.ndarray, world: int) -> np.ndarray:
    """'same' mode computed slab by slab: d[g] = sum_t k[t] a[g + c0 - t] needs K0-1-c0 planes from below
    and c0 = K0//2 planes from above (zeros beyond the volume ends); the interior sits at full-convolution
    index K0-1 of the local problem.  Must equal conv_same(activity, kernel)."""
    a = np.asarray(activity, dtype=np.float64)
    k = np.asarray(kernel, dtype=np.float64)
    n0 = a.shape[0]
    K0 = k.shape[0]
    c0 = K0 // 2
    dn = K0 - 1 - c0
    c1, c2 = k.shape[1] // 2, k.shape[2] // 2
    out = np.empty_like(a)
    for lo, hi in slab_bounds(n0, world):
        local = np.zeros((hi - lo + K0 - 1,) + a.shape[1:], dtype=np.float64)
        src_lo, src_hi = max(lo - dn, 0), min(hi + c0, n0)
        local[src_lo - (lo - dn) : src_hi - (lo - dn)] = a[src_lo:src_hi]
        big = tuple(n + kk - 1 for n, kk in zip(local.shape, k.shape))
        pad = np.zeros(big)
        pad[: local.shape[0], : local.shape[1], : local.shape[2]] = local
        full = conv_reference(pad, k)
        out[lo:hi] = full[K0 - 1 : K0 - 1 + (hi - lo), c1 : c1 + a.shape[1], c2 : c2 + a.shape[2]]
    return out


# --------------------------------------------------------------------------------------
# Section 8f "next" rows: the steps either side of the convolution
# --------------------------------------------------------------------------------------


def fit_monoexp_curvefit(times, activities, half_life: float, weight_factors=None):
    """Literal restatement of TimeCurveFitting.fit_time_activity_curve
    (time_integration/curve_fitting.py:36-65): one scipy.optimize.curve_fit call per voxel with
    p0 = [y(t_0), ln2/half_life], sigma = 1/weight_factors; a raised exception gives [0, decay_constant];
    accumulated = A0/lambda*(1 - exp(-lambda*100*half_life)) (:74-84).  The arithmetic lives in SciPy
    (MINPACK lmdif through scipy.optimize.leastsq; the reference pins only scipy>=1.7.0, setup.py)."""
    import warnings

    from scipy.optimize import curve_fit

    decay_constant = np.log(2) / half_life
    times = np.array(times)
    activities = np.array(activities)
    if weight_factors is None:
        weight_factors = np.ones_like(times)
    n_times = len(times)
    original_shape = activities[0].shape
    flat = activities.reshape(n_times, -1)
    fitted = np.zeros((2, flat.shape[1]))

    def decay(t, A0, lam):
        return A0 * np.exp(-lam * t)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(flat.shape[1]):
            try:
                popt, _ = curve_fit(decay, times, flat[:, i], p0=[flat[0, i], decay_constant], sigma=1 / np.array(weight_factors))
                fitted[:, i] = popt
            except Exception:
                fitted[:, i] = [0, decay_constant]
    integration_limit = 100 * half_life  # curve_fitting.py:78-79
    acc = fitted[0] / fitted[1] * (1 - np.exp(-fitted[1] * integration_limit))
    return fitted.reshape((2, *original_shape)), acc.reshape(original_shape)


TISSUE_HU_RANGES = {  # tissue/composition.py:40-46
    "air": (-1000, -900), "lung": (-900, -500), "soft_tissue": (-100, 100), "bone": (300, 3000), "water": (-10, 10),
}


def handle_artifacts(ct_image: np.ndarray) -> np.ndarray:
    """tissue/composition.py:73-93: metal (> 2000 HU) voxels replaced by gaussian_filter(sigma=1) of the image
    with those voxels zeroed (NaN -> nan_to_num -> 0)."""
    from scipy.ndimage import gaussian_filter

    ct_image = np.asarray(ct_image, dtype=np.float64)
    metal_mask = ct_image > 2000
    corrected = ct_image.copy()
    corrected[metal_mask] = np.nan
    smoothed = gaussian_filter(np.nan_to_num(corrected), sigma=1)
    corrected[metal_mask] = smoothed[metal_mask]
    return corrected


def tissue_composition(ct_image: np.ndarray, handle: bool = True) -> dict:
    """tissue/composition.py:48-71: dict of 0/1 float maps per tissue class by inclusive HU range."""
    ct = handle_artifacts(ct_image) if handle else np.asarray(ct_image, dtype=np.float64)
    return {name: ((ct >= lo) & (ct <= hi)).astype(float) for name, (lo, hi) in TISSUE_HU_RANGES.items()}


def calculate_dvh(dose_map: np.ndarray, roi_mask: np.ndarray, bins: int = 1000):
    """core/utils.py:233-262, literally: np.histogram over the ROI doses, cumulative volume fraction."""
    if dose_map.shape != roi_mask.shape:
        raise ValueError("Dose map and ROI mask must have same dimensions")
    roi_doses = dose_map[roi_mask > 0]
    if len(roi_doses) == 0:
        raise ValueError("ROI mask is empty")
    hist, edges = np.histogram(roi_doses, bins=bins)
    cum_dvh = 1.0 - np.cumsum(hist) / len(roi_doses)
    return edges[1:], cum_dvh


def load_kernel_dat(filename):
    """core/utils.py:17-51, literally (np.fromfile field by field; 80-byte header)."""
    from datetime import datetime

    with open(filename, "rb") as f:
        dims = np.fromfile(f, dtype=np.int32, count=3)
        voxel_size = np.fromfile(f, dtype=np.float32, count=1)[0]
        total_energy = np.fromfile(f, dtype=np.float32, count=1)[0]
        scaling = np.fromfile(f, dtype=np.float32, count=1)[0]
        timestamp = np.fromfile(f, dtype=np.int32, count=6)
        user = np.fromfile(f, dtype=np.int8, count=32)
        kernel = np.fromfile(f, dtype=np.float32)
        kernel = kernel.reshape(dims) * scaling
    metadata = {
        "dimensions": dims, "voxel_size": voxel_size, "total_energy": total_energy, "scaling_factor": scaling,
        "creation_date": datetime(*timestamp).strftime("%Y-%m-%d %H:%M:%S"),
        "created_by": bytes(user).decode().strip("\x00"),
    }
    return kernel, metadata


# --------------------------------------------------------------------------------------
# Synthetic inputs of SURVEY section 8d (fixed seeds) - shared by tests and bench.py
# --------------------------------------------------------------------------------------


# --------------------------------------------------------------------------------------
# interpolate_timepoints   core/utils.py:154-191
# The arithmetic lives in scipy.interpolate.interp1d (third party, not vendored; the reference pins
# `scipy>=1.6.0`, here scipy 1.18.1).  Restated from its published algorithm: sort by x (stable), then
#   linear    y = (x_new - x_lo)/(x_hi - x_lo) * y_hi + (x_hi - x_new)/(x_hi - x_lo) * y_lo on the bracketing pair
#             (searchsorted, clipped to [1, n-1]: the end segments extrapolate);
#   previous  the last sample at or before x_new; NaN before the first sample, the last sample beyond the end
#             (what interp1d does for fill_value='extrapolate');
#   cubic     the not-a-knot cubic spline through all samples (make_interp_spline(k=3) default), extended
#             beyond the ends by its end polynomials.
# Pinned by oracle/gen_golden.py against the real reference function (tests/golden/interp_ref.npz): linear and
# previous bit for bit, cubic to 1e-12 relative (a different but equivalent linear solve).
# --------------------------------------------------------------------------------------


def _notaknot_second_derivatives(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Second derivatives M_i at the samples of the C2 piecewise cubic through (x_i, y_i[...]) whose third derivative is
    continuous at x_1 and x_{n-2} (not-a-knot).  y has shape (n, ...)."""
    n = len(x)
    h = np.diff(x)
    A = np.zeros((n, n))
    rhs = np.zeros((n,) + y.shape[1:])
    for i in range(1, n - 1):  # continuity of the first derivative at the interior samples
        A[i, i - 1], A[i, i], A[i, i + 1] = h[i - 1] / 6, (h[i - 1] + h[i]) / 3, h[i] / 6
        rhs[i] = (y[i + 1] - y[i]) / h[i] - (y[i] - y[i - 1]) / h[i - 1]
    # third derivative (M_{i+1} - M_i)/h_i continuous at x_1 and x_{n-2}
    A[0, 0], A[0, 1], A[0, 2] = h[1], -(h[0] + h[1]), h[0]
    A[n - 1, n - 3], A[n - 1, n - 2], A[n - 1, n - 1] = h[n - 2], -(h[n - 3] + h[n - 2]), h[n - 3]
    return np.linalg.solve(A, rhs.reshape(n, -1)).reshape(rhs.shape)


def interpolate_timepoints(time_points: Sequence[float], values: Sequence[np.ndarray], new_times: Sequence[float],
                           method: str = "linear") -> List[np.ndarray]:
    """core/utils.py:154-191: the sampled 3-D arrays interpolated voxel by voxel along time."""
    if len(time_points) != len(values):
        raise ValueError("Number of time points must match number of values")
    shape = np.asarray(values[0]).shape
    x = np.asarray(time_points, dtype=np.float64)
    y = np.array([np.asarray(v, dtype=np.float64).reshape(-1) for v in values])
    order = np.argsort(x, kind="mergesort")
    x, y = x[order], y[order]
    xn = np.asarray(new_times, dtype=np.float64)
    n = len(x)
    if method == "linear":
        hi = np.searchsorted(x, xn).clip(1, n - 1).astype(int)
        lo = hi - 1
        out = ((xn - x[lo]) / (x[hi] - x[lo]))[:, None] * y[hi] + ((x[hi] - xn) / (x[hi] - x[lo]))[:, None] * y[lo]
    elif method == "previous":
        idx = np.searchsorted(np.nextafter(x, -np.inf), xn, side="left").clip(1, n).astype(int)
        out = y[idx - 1].copy()
        out[xn < x[0]] = np.nan
        out[xn > x[-1]] = y[-1]
    elif method == "cubic":
        if n < 4:
            raise ValueError("x and y arrays must have at least 4 entries")
        M = _notaknot_second_derivatives(x, y)
        seg = (np.searchsorted(x, xn, side="right") - 1).clip(0, n - 2)  # beyond the ends: the end polynomials
        h = (x[seg + 1] - x[seg])[:, None]
        a, b = (x[seg + 1] - xn)[:, None], (xn - x[seg])[:, None]
        out = (M[seg] * a ** 3 + M[seg + 1] * b ** 3) / (6 * h) + (y[seg] / h - M[seg] * h / 6) * a + (y[seg + 1] / h - M[seg + 1] * h / 6) * b
    else:
        raise NotImplementedError(f"{method} is not restated")
    return [v.reshape(shape) for v in out]


def sphere_activity(size=(48, 48, 48), center=(24, 24, 24), radius=8, activity=2e6) -> np.ndarray:
    """examples/single_timepoint_y90_physical_decay.py:14-21 (vectorised, same voxels)."""
    x, y, z = np.ogrid[: size[0], : size[1], : size[2]]
    m = (x - center[0]) ** 2 + (y - center[1]) ** 2 + (z - center[2]) ** 2 <= radius**2
    arr = np.zeros(size)
    arr[m] = activity
    return arr


def rel_err_of_peak(got: np.ndarray, ref: np.ndarray) -> float:
    """The north-star parity metric: max|got - ref| / max|ref|."""
    ref = np.asarray(ref, dtype=np.float64)
    denom = float(np.max(np.abs(ref)))
    if denom == 0.0:
        return float(np.max(np.abs(np.asarray(got, dtype=np.float64))))
    return float(np.max(np.abs(np.asarray(got, dtype=np.float64) - ref)) / denom)
