#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE (/root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Recipe:
  1. copy /root/reference/pyvoxeldosimetry to a writable temp dir (the factory mkdirs and
     np.saves inside the package, data/dose_kernels/kernel_factory.py:34-35,66-73);
  2. register permissive stub modules for the absent third-party imports (nibabel, SimpleITK,
     pydicom, cupy, matplotlib) - the hot path never touches them;
  3. import the real KernelConvolutionCalculator / generators / ActivitySampler and run them.
For Lu177 the JSON `//` comments (Lu177/Lu177.json:16,21,30,32) are stripped before json.loads.

The script also asserts that oracle/dose_oracle.py reproduces every generated vector (bit-exact
for the FFT expression, <=1e-15 relative for the generators), i.e. it PINS the oracle.
The GPU box has no /root/reference: it only reads the committed .npz files.

Usage:  python oracle/gen_golden.py            (writes tests/golden/)
"""
from __future__ import annotations

import json
import os
import re
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")


def import_reference():
    """The real reference, copied to oracle/_ref and imported under the stub recipe (oracle/ref_loader.py)."""
    sys.path.insert(0, REPO)
    from oracle import ref_loader

    ref_loader.build_ref()
    ref_loader.import_reference()
    return None


def gen_next_rows(orc):
    """SURVEY section 8f rows run through the REAL reference: per-voxel curve fit, CT artifact handling + tissue
    classes, DVH, `.dat` kernel reader.  Own rng so the vectors above never change."""
    import warnings

    from pyvoxeldosimetry.time_integration.curve_fitting import TimeCurveFitting
    from pyvoxeldosimetry.tissue.composition import TissueComposition
    from pyvoxeldosimetry.core import utils as ref_utils

    rng = np.random.default_rng(8061)
    out = {}
    # ---- mono-exponential fit (curve_fitting.py:19-65): noisy decaying curves, unweighted and weighted
    hl = 161.52
    lam0 = np.log(2) / hl
    for name, times, shape, noise, weights in (("fit4", [4.0, 24.0, 96.0, 168.0], (5, 4, 6), 0.05, None),
                                               ("fit3w", [2.0, 20.0, 70.0], (3, 4, 5), 0.10, [1.0, 2.0, 0.5]),
                                               ("fit6", [1.0, 4.0, 24.0, 48.0, 96.0, 168.0], (2, 3, 4), 0.15, None)):
        A0 = rng.uniform(1e2, 1e6, shape)
        lam = lam0 * rng.uniform(0.7, 4.0, shape)
        maps = [A0 * np.exp(-lam * t) * (1 + noise * rng.standard_normal(shape)) for t in times]
        maps[-1][0, 0, 0] = maps[-2][0, 0, 0] = 0.0  # a voxel that has decayed to exactly zero
        for m in maps:
            m[0, 0, 1] = 0.0                         # an all-zero voxel
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            params, acc = TimeCurveFitting(hl).fit_time_activity_curve(times, maps, weights)
        p2, a2 = orc.fit_monoexp_curvefit(times, maps, hl, weights)
        assert np.array_equal(params, p2) and np.array_equal(acc, a2, equal_nan=True), name
        out[name + "|times"], out[name + "|maps"], out[name + "|params"], out[name + "|acc"] = np.array(times), np.stack(maps), params, acc
        if weights is not None:
            out[name + "|weights"] = np.array(weights)
    # ---- CT: artifact handling + tissue classes (composition.py:48-93)
    ct = rng.uniform(-1000, 1500, (14, 12, 10))
    ct[3:6, 4:7, 2:5] = rng.uniform(2100, 3500, (3, 3, 3))  # an implant
    ct[0, 0, 0] = 2600.0                                      # metal on a corner: exercises 'reflect'
    ct[13, 11, 9] = 2001.0
    ct[7, 7, 7], ct[7, 7, 8], ct[8, 8, 8] = -900.0, 100.0, 2000.0  # inclusive range edges / threshold itself
    tc = TissueComposition()
    corr = tc._handle_artifacts(ct)
    comp = tc.calculate_composition(ct, handle_artifacts=True)
    comp_raw = tc.calculate_composition(ct, handle_artifacts=False)
    assert np.array_equal(orc.handle_artifacts(ct), corr)
    for k, v in orc.tissue_composition(ct, True).items():
        assert np.array_equal(v, comp[k]), k
    out["ct|hu"], out["ct|corrected"] = ct, corr
    for k in comp:
        out[f"ct|comp|{k}"] = comp[k].astype(np.uint8)
        out[f"ct|comp_raw|{k}"] = comp_raw[k].astype(np.uint8)
    # ---- DVH (core/utils.py:233-262) on a float32 dose map (the dose path's output type)
    dose = (rng.gamma(2.0, 5.0, (9, 8, 7)) * 1e3).astype(np.float32)
    mask = rng.uniform(0, 1, dose.shape) > 0.4
    for bins in (1000, 17):
        e, c = ref_utils.calculate_dvh(dose, mask, bins)
        e2, c2 = orc.calculate_dvh(dose, mask, bins)
        assert np.array_equal(e, e2) and np.array_equal(c, c2)
        out[f"dvh|edges{bins}"], out[f"dvh|cum{bins}"] = e, c
    out["dvh|dose"], out["dvh|mask"] = dose, mask.astype(np.uint8)
    # ---- `.dat` kernel file: bytes laid out per utils.py:31-41, parsed by the reference reader
    k = rng.uniform(0, 1, (4, 5, 6)).astype(np.float32)
    import struct
    blob = struct.pack("<3i3f6i", 4, 5, 6, 1.5, 0.9337, 2.0, 2025, 2, 8, 9, 50, 56) + b"devhliu".ljust(32, b"\x00") + k.tobytes()
    path = os.path.join(tempfile.mkdtemp(prefix="pvd_dat_"), "k.dat")
    with open(path, "wb") as f:
        f.write(blob)
    kr, md = ref_utils.load_kernel(path)
    ko, mo = orc.load_kernel_dat(path)
    assert np.array_equal(kr, ko) and md["creation_date"] == mo["creation_date"] == "2025-02-08 09:50:56" and md["created_by"] == "devhliu"
    out["dat|blob"], out["dat|kernel"] = np.frombuffer(blob, dtype=np.uint8), kr
    np.savez_compressed(os.path.join(OUT, "next_ref.npz"), **out)


def gen_interp(orc):
    """interpolate_timepoints (core/utils.py:154-191) run through the REAL reference (scipy interp1d underneath): unsorted
    and non-uniform time points, new times inside, on and beyond the sampled range, the three kinds the docstring names."""
    from pyvoxeldosimetry.core import utils as ref_utils

    rng = np.random.default_rng(154191)
    out = {}
    for name, times, shape in (("s5", [24.0, 4.0, 96.0, 168.0, 48.0], (4, 3, 5)), ("s4", [1.0, 3.0, 7.5, 20.0], (2, 3, 4)),
                               ("s7", [0.0, 1.0, 2.0, 4.0, 8.0, 16.0, 32.0], (3, 2, 2))):
        vals = [rng.uniform(0.0, 1e4, shape) * np.exp(-0.01 * t) for t in times]
        new = [min(times) - 3.0, times[0], 5.5, 30.0, max(times), max(times) + 50.0, min(times), 6.0]
        out[name + "|times"], out[name + "|values"], out[name + "|new"] = np.array(times), np.stack(vals), np.array(new)
        for method in ("linear", "cubic", "previous"):
            ref = np.stack(ref_utils.interpolate_timepoints(times, vals, new, method))
            mine = np.stack(orc.interpolate_timepoints(times, vals, new, method))
            if method == "cubic":
                assert np.allclose(mine, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max()), (name, method)
            else:
                assert np.array_equal(mine, ref, equal_nan=True), (name, method)
            out[f"{name}|{method}"] = ref
    np.savez_compressed(os.path.join(OUT, "interp_ref.npz"), **out)


def main():
    tmp = import_reference()
    sys.path.insert(0, REPO)
    from oracle import dose_oracle as orc
    from pyvoxeldosimetry.core.kernel_convolution import KernelConvolutionCalculator
    from pyvoxeldosimetry.core.activity_sampler import ActivitySampler
    from pyvoxeldosimetry.data.dose_kernels.y90_kernel import Y90KernelGenerator
    from pyvoxeldosimetry.data.dose_kernels.lu177_kernel import Lu177KernelGenerator
    from pyvoxeldosimetry.time_integration.curve_fitting import TimeCurveFitting

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)

    # ---------------- kernels (A5 / A6) ----------------
    kern = {}
    tissues = ["water", "lung", "soft_tissue", "bone", "iodine_contrast", "no_such_tissue"]
    with np.errstate(divide="ignore", invalid="ignore"):
        for t in tissues:
            for vox, grid in [(1.0, (15, 15, 15)), (2.0, (12, 10, 8)), (4.8, (9, 9, 9))]:
                k = Y90KernelGenerator(t).generate_kernel(vox, grid)
                key = f"Y90|{t}|{vox}|{'x'.join(map(str, grid))}"
                kern[key] = k
                c = tuple(g // 2 for g in grid)
                assert np.isnan(k[c]) and np.isnan(k).sum() == 1, "reference Y90 centre is expected to be the only NaN"
                mine = orc.y90_kernel(vox, grid, t)
                m = np.ones(grid, bool)
                m[c] = False
                assert np.allclose(mine[m], k[m], rtol=1e-14, atol=0), key
                lit = orc.y90_kernel(vox, grid, t, centre="reference")
                assert np.isnan(lit[c])
            for vox, grid in [(1.0, (15, 15, 15)), (4.8, (11, 11, 11)), (2.5, (8, 9, 10))]:
                k = Lu177KernelGenerator(t).generate_kernel(vox, grid)
                key = f"Lu177|{t}|{vox}|{'x'.join(map(str, grid))}"
                kern[key] = k
                assert np.isfinite(k).all()
                assert np.allclose(orc.lu177_kernel(vox, grid, t), k, rtol=1e-14, atol=0), key
    # scalar known answers on the production grids
    with np.errstate(divide="ignore", invalid="ignore"):
        k64 = Y90KernelGenerator("water").generate_kernel(1.0, (64, 64, 64))
    kats = {
        "y90_water_1mm_64_c+1": k64[33, 32, 32],
        "y90_water_2mm_64_c+1": Y90KernelGenerator("water").generate_kernel(2.0, (64, 64, 64))[33, 32, 32],
        "y90_lung_1mm_64_c+1": Y90KernelGenerator("lung").generate_kernel(1.0, (64, 64, 64))[33, 32, 32],
        "y90_bone_1mm_64_c+1": Y90KernelGenerator("bone").generate_kernel(1.0, (64, 64, 64))[33, 32, 32],
    }
    l31 = Lu177KernelGenerator("water").generate_kernel(4.8, (31, 31, 31))
    l81 = Lu177KernelGenerator("water").generate_kernel(1.0, (81, 81, 81))
    kats.update({
        "lu177_water_4.8mm_31_centre": l31[15, 15, 15], "lu177_water_4.8mm_31_c+1": l31[16, 15, 15],
        "lu177_water_4.8mm_31_sum": l31.sum(), "lu177_water_1mm_81_c+1": l81[41, 40, 40], "lu177_water_1mm_81_sum": l81.sum(),
    })
    np.savez_compressed(os.path.join(OUT, "kernels_ref.npz"), **kern)

    # ---------------- convolution through the real calculator (A1, A2) ----------------
    calc = KernelConvolutionCalculator("Y90", "water", 1.0)  # builds (NaN-centred) 64^3 kernel
    assert calc.kernel.shape == (64, 64, 64)
    conv = {}
    cases = {
        "pad_16x12x20_k5x7x3": ((16, 12, 20), (5, 7, 3)),
        "crop_10x9x8_k12x4x11": ((10, 9, 8), (12, 4, 11)),       # kernel larger than the grid on 2 axes
        "odd_7x11x13_k3x3x3": ((7, 11, 13), (3, 3, 3)),          # prime lengths
        "k1_8x8x8_k1x1x1": ((8, 8, 8), (1, 1, 1)),
        "even_k_12x10x14_k4x6x2": ((12, 10, 14), (4, 6, 2)),
    }
    for name, (ashape, kshape) in cases.items():
        a = rng.uniform(0.0, 1e3, size=ashape)
        a[tuple(s // 2 for s in ashape)] = 2e6
        k = rng.uniform(0.0, 1.0, size=kshape)
        calc.kernel = k
        d = calc.calculate_dose_rate(a, (1.0, 1.0, 1.0))
        conv[name + "|a"], conv[name + "|k"], conv[name + "|d"] = a, k, d
        assert np.array_equal(orc.conv_reference(a, k), d), name           # bit-exact restatement
        bf = orc.conv_bruteforce(a, k)
        assert orc.rel_err_of_peak(bf, d) < 1e-13, name                    # definition check
    # multi-timepoint (A2) + activity trapezoid (A3)
    ashape, kshape = (12, 10, 14), (5, 5, 5)
    times = [4.0, 24.0, 96.0, 168.0]
    a0 = rng.uniform(0.0, 1e3, size=ashape)
    maps = [a0 * np.exp(-np.log(2) * t / 161.52) for t in times]
    k = rng.uniform(0.0, 1.0, size=kshape)
    calc.kernel = k
    D = calc.calculate_absorbed_dose(maps, times, (1.0, 1.0, 1.0))
    assert np.array_equal(orc.absorbed_dose_trapezoid(maps, times, k), D)
    w = orc.trapezoid_weights(times, 3600.0)
    assert orc.rel_err_of_peak(orc.conv_reference(sum(wi * m for wi, m in zip(w, maps)), k), D) < 1e-14
    A = ActivitySampler(161.52).integrate_activity(maps, times)
    assert np.array_equal(orc.integrate_activity_trapezoid(maps, times), A)
    conv["tp|maps"], conv["tp|times"], conv["tp|k"], conv["tp|D"], conv["tp|A"] = np.stack(maps), np.array(times), k, D, A
    # A11 closed-form integral
    tcf = TimeCurveFitting(161.52)
    params = np.stack([rng.uniform(0, 1e4, 50), rng.uniform(1e-3, 1e-1, 50)])
    acc = tcf._calculate_accumulated_dose(params)
    assert np.array_equal(orc.accumulated_activity_monoexp(params[0], params[1], 161.52), acc)
    acc2 = tcf._calculate_accumulated_dose(params, integration_limit=72.0)
    conv["a11|params"], conv["a11|acc"], conv["a11|acc72"] = params, acc, acc2
    np.savez_compressed(os.path.join(OUT, "conv_ref.npz"), **conv)

    # ---------------- config C1: the Y90 example (single_timepoint_y90_physical_decay.py) ----------------
    sphere = orc.sphere_activity()
    # literal reference: all-NaN because of the NaN centre voxel
    calc2 = KernelConvolutionCalculator("Y90", "water", 1.0)
    calc2.kernel = k64.copy()
    with np.errstate(invalid="ignore"):
        d_nan = calc2.calculate_dose_rate(sphere, (1.0, 1.0, 1.0))
    kats["c1_literal_nan_voxels"] = float(np.isnan(d_nan).sum())
    kfin = k64.copy()
    kfin[32, 32, 32] = 1.0  # beta(0)*rho*S*f = 1 for water; brems(0) := 0
    assert np.array_equal(orc.y90_kernel(1.0, (64, 64, 64), "water")[32, 32, 32], 1.0)
    calc2.kernel = kfin
    d1 = calc2.calculate_dose_rate(sphere, (1.0, 1.0, 1.0))
    assert np.array_equal(orc.conv_reference(sphere, orc.y90_kernel(1.0, (64, 64, 64))), orc.conv_reference(sphere, kfin)) or \
        orc.rel_err_of_peak(orc.conv_reference(sphere, orc.y90_kernel(1.0, (64, 64, 64))), d1) < 1e-14
    kats.update({
        "c1_sphere_voxels": float((sphere > 0).sum()), "c1_sum_a": sphere.sum(), "c1_max": d1.max(),
        "c1_argmax": float(np.ravel_multi_index(np.unravel_index(d1.argmax(), d1.shape), d1.shape)),
        "c1_sum": d1.sum(), "c1_d000": d1[0, 0, 0], "c1_d242424": d1[24, 24, 24], "c1_min": d1.min(),
    })
    # keep three orthogonal central planes through the peak (8,8,8) as array fixtures (small)
    np.savez_compressed(os.path.join(OUT, "c1_ref.npz"), plane_x8=d1[8], plane_y8=d1[:, 8], plane_z8=d1[:, :, 8])
    gen_next_rows(orc)
    gen_interp(orc)
    with open(os.path.join(OUT, "kats.json"), "w") as f:
        json.dump({k: float(v) for k, v in kats.items()}, f, indent=1, sort_keys=True)
    print("golden written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(f"  {fn}: {os.path.getsize(os.path.join(OUT, fn))} bytes")


if __name__ == "__main__":
    main()
