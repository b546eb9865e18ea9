"""interpolate_timepoints (reference core/utils.py:154-191), the step before the time integration: the sampled activity
volumes interpolated voxel by voxel along time.  scipy's interp1d is a linear map of the samples, so the product derives a
J x T weight matrix on the host and runs ONE pass over the volumes (pvd_weighted_combine).

CPU part: the oracle restatement and the product's weights against vectors the REAL reference function produced
(tests/golden/interp_ref.npz, oracle/gen_golden.py), the combine kernel through the SIMT emulator.  GPU part: the product
API against the golden vectors and the oracle.  Tolerance: float32 arithmetic on float32 volumes - 2e-6 of the peak of the
interpolated series (the weights themselves are float64-exact to 1e-12).
"""
import os

import numpy as np
import pytest

from oracle import dose_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ("s5", "s4", "s7")
KINDS = ("linear", "cubic", "previous")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "interp_ref.npz"))


def _apply(W, vals):
    """float64 application of a weight matrix with the kernel's conventions (zero weight skips, NaN row -> NaN)."""
    T = len(vals)
    flat = np.stack([np.asarray(v, dtype=np.float64).reshape(-1) for v in vals])
    out = []
    for row in W:
        acc = np.zeros(flat.shape[1])
        for t in range(T):
            if row[t] != 0:
                acc = acc + row[t] * flat[t]
        out.append(acc.reshape(np.asarray(vals[0]).shape))
    return out


def test_oracle_matches_the_reference_vectors(gold):
    for name in CASES:
        times, vals, new = gold[name + "|times"].tolist(), list(gold[name + "|values"]), gold[name + "|new"].tolist()
        for kind in KINDS:
            ref = gold[f"{name}|{kind}"]
            mine = np.stack(orc.interpolate_timepoints(times, vals, new, kind))
            if kind == "cubic":
                np.testing.assert_allclose(mine, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
            else:
                assert np.array_equal(mine, ref, equal_nan=True)
            assert np.isnan(ref).any() == (kind == "previous")  # 'previous' before the first sample is NaN in the reference
    with pytest.raises(ValueError):
        orc.interpolate_timepoints([1.0, 2.0], [np.zeros((2, 2, 2))], [1.5])


def test_product_weights_reproduce_the_reference_vectors(gold):
    from pyvoxeldosimetry_b200.core.utils import interpolation_weights

    for name in CASES:
        times, vals, new = gold[name + "|times"].tolist(), list(gold[name + "|values"]), gold[name + "|new"].tolist()
        for kind in KINDS:
            ref = gold[f"{name}|{kind}"]
            W = interpolation_weights(times, new, kind)
            assert W.shape == (len(new), len(times)) and W.dtype == np.float64
            got = np.stack(_apply(W, vals))
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12 * np.nanmax(np.abs(ref)), equal_nan=True)
            if kind == "linear":
                assert ((W != 0).sum(axis=1) <= 2).all()  # two samples per interpolated volume: the others are never loaded
    # the other step kinds of interp1d that the reference's `kind=method` pass-through reaches
    from scipy.interpolate import interp1d

    rng = np.random.default_rng(5)
    times, new = [9.0, 1.0, 4.0, 2.5], [0.0, 1.0, 1.7, 3.25, 4.0, 9.0, 11.0]
    y = rng.uniform(0, 1, (4, 30))
    for kind in ("next", "nearest", "slinear"):
        ref = interp1d(times, y, axis=0, kind=kind, bounds_error=False, fill_value="extrapolate")(new)
        got = np.stack(_apply(interpolation_weights(times, new, kind), list(y)))
        got[np.isnan(interpolation_weights(times, new, kind)).any(axis=1)] = np.nan
        np.testing.assert_allclose(got, ref, rtol=1e-12, equal_nan=True)
    with pytest.raises(ValueError):
        interpolation_weights([1.0, 2.0, 3.0], [1.5], "cubic")
    with pytest.raises(NotImplementedError):
        interpolation_weights([1.0, 2.0, 3.0], [1.5], "quintic")


def test_emulated_combine_kernel(gold):
    """The real weighted_combine_kernel through the SIMT emulator: every J template, ragged tails, unaligned pointers
    (scalar path), unused volumes that hold NaN, a NaN weight row, an output that is one of the inputs."""
    import emu_util

    lib = emu_util.emu_lib()
    rng = np.random.default_rng(11)
    for T, J, n, shift in ((1, 1, 64, 0), (5, 1, 777, 0), (4, 3, 1001, 0), (7, 8, 130, 1), (16, 16, 259, 0), (3, 9, 5, 3)):
        raw = [rng.uniform(-1, 1, n + 4).astype(np.float32) for _ in range(T)]
        vols = [r[shift:shift + n] for r in raw]
        W = rng.uniform(-2, 2, (J, T))
        W[rng.uniform(0, 1, (J, T)) < 0.3] = 0.0
        if T >= 3:
            W[:, 1] = 0.0
            vols[1][:] = np.nan  # never referenced: must not poison the sums
        outs_raw = [np.full(n + 4, 7.0, np.float32) for _ in range(J)]
        outs = [o[shift:shift + n] for o in outs_raw]
        lib.weighted_combine([v.ctypes.data for v in vols], W.tolist(), [o.ctypes.data for o in outs], n)
        W32 = W.astype(np.float32).astype(np.float64)
        for j in range(J):
            ref = _apply(W32[j:j + 1], vols)[0]
            np.testing.assert_allclose(outs[j], ref, rtol=0, atol=2e-6 * max(1.0, np.abs(ref).max()))
            assert (outs_raw[j][:shift] == 7.0).all() and (outs_raw[j][shift + n:] == 7.0).all()  # nothing written outside
    v = [rng.uniform(0, 1, 100).astype(np.float32) for _ in range(2)]
    o = np.empty(100, np.float32)
    lib.weighted_combine([x.ctypes.data for x in v], [[float("nan"), 0.0]], [o.ctypes.data], 100)
    assert np.isnan(o).all()
    keep = v[0].copy()
    lib.weighted_combine([v[0].ctypes.data, v[1].ctypes.data], [[1.0, 2.0]], [v[0].ctypes.data], 100)  # in place
    np.testing.assert_allclose(v[0], keep + 2.0 * v[1], rtol=1e-6)
    # the reference vectors through the kernel
    from pyvoxeldosimetry_b200.core.utils import interpolation_weights

    for name in CASES:
        times, new = gold[name + "|times"].tolist(), gold[name + "|new"].tolist()
        vals = [np.ascontiguousarray(x, dtype=np.float32) for x in gold[name + "|values"]]
        for kind in KINDS:
            ref = gold[f"{name}|{kind}"]
            outs = [np.empty(vals[0].shape, np.float32) for _ in new]
            lib.weighted_combine([x.ctypes.data for x in vals], interpolation_weights(times, new, kind).tolist(),
                                 [o.ctypes.data for o in outs], vals[0].size)
            got = np.stack(outs)
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            assert np.nanmax(np.abs(got - ref)) <= 2e-6 * np.nanmax(np.abs(ref))


@pytest.mark.gpu
def test_gpu_interpolate_timepoints_vs_reference_vectors_and_oracle(gold):
    import torch
    from pyvoxeldosimetry_b200.core.utils import interpolate_timepoints
    from pyvoxeldosimetry_b200 import engine

    for name in CASES:
        times, vals, new = gold[name + "|times"].tolist(), list(gold[name + "|values"]), gold[name + "|new"].tolist()
        for kind in KINDS:
            ref = gold[f"{name}|{kind}"]
            got = interpolate_timepoints(times, vals, new, kind)
            assert isinstance(got, list) and len(got) == len(new) and got[0].shape == vals[0].shape
            got = np.stack(got)
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            assert np.nanmax(np.abs(got - ref)) <= 2e-6 * np.nanmax(np.abs(ref))
    # a C2-like series on the device: 4 samples of 96 x 80 x 72 voxels (odd total: ragged tail), 11 new times incl. extrapolation
    rng = np.random.default_rng(177)
    shape, times = (96, 80, 73), [4.0, 24.0, 96.0, 168.0]
    a0 = rng.uniform(0, 1e3, shape)
    maps = [a0 * np.exp(-np.log(2) * t / 161.52) * (1 + 0.02 * rng.standard_normal(shape)) for t in times]
    new = list(np.linspace(0.0, 200.0, 11))
    dev = [torch.from_numpy(m.astype(np.float32)).cuda() for m in maps]
    maps32 = [m.astype(np.float32) for m in maps]
    for kind in ("linear", "cubic", "previous"):
        got = interpolate_timepoints(times, dev, new, kind)
        assert all(g.is_cuda for g in got)
        ref = np.stack(orc.interpolate_timepoints(times, maps32, new, kind))
        got = torch.stack(got).cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert np.nanmax(np.abs(got - ref)) <= 2e-6 * np.nanmax(np.abs(ref))
    # interpolate -> trapezoid -> convolution == the oracle pipeline (the interpolated series feeds the dose path)
    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator

    calc = KernelConvolutionCalculator("Lu177", "water", 4.8, config={"kernel_grid": (9, 9, 9)})
    fine = interpolate_timepoints(times, dev, new[:9], "linear")
    dose = calc.calculate_absorbed_dose(fine, new[:9], (4.8, 4.8, 4.8))
    dose = dose.cpu().numpy() if hasattr(dose, "cpu") else dose
    ref_series = orc.interpolate_timepoints(times, maps32, new[:9], "linear")
    ref = orc.absorbed_dose_trapezoid(ref_series, new[:9], calc.kernel.astype(np.float32).astype(np.float64))
    assert orc.rel_err_of_peak(dose, ref) <= 1e-4
    # more outputs than one pass holds (17 > 16) and more samples than one pass reads (20 > 16)
    many_new = list(np.linspace(5.0, 160.0, 17))
    got = torch.stack(interpolate_timepoints(times, dev, many_new, "linear")).cpu().numpy()
    ref = np.stack(orc.interpolate_timepoints(times, maps32, many_new, "linear"))
    assert np.max(np.abs(got - ref)) <= 2e-6 * np.abs(ref).max()
    t20 = list(np.linspace(0.0, 190.0, 20))
    series = [torch.from_numpy((a0 * np.exp(-0.004 * t)).astype(np.float32)).cuda() for t in t20]
    w = rng.uniform(-1, 1, (2, 20))
    got = engine.weighted_combine(series, w.tolist())
    for j in range(2):
        ref = sum(float(np.float32(w[j, t])) * series[t].double() for t in range(20))
        mag = sum(abs(float(w[j, t])) * series[t].double() for t in range(20))  # the signed weights cancel: scale by the terms
        assert float((got[j].double() - ref).abs().max()) <= 2e-6 * float(mag.max())
    with pytest.raises(ValueError):
        interpolate_timepoints([1.0, 2.0], [maps[0]], [1.5])
