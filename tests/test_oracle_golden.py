"""The oracle against the golden vectors produced by the REAL reference (oracle/gen_golden.py) and against
FFT-free definitions.  CPU only."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import dose_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_kernel_generators_match_reference():
    z = np.load(os.path.join(GOLD, "kernels_ref.npz"))
    assert len(z.files) == 36
    for key in z.files:
        nuc, tissue, vox, grid = key.split("|")
        grid = tuple(int(g) for g in grid.split("x"))
        ref = z[key]
        mine = orc.make_kernel(nuc, float(vox), grid, tissue)
        c = tuple(g // 2 for g in grid)
        if nuc == "Y90":  # the reference has exactly one NaN, at r = 0 (y90_kernel.py:134-138)
            assert np.isnan(ref[c]) and np.isnan(ref).sum() == 1
            lit = orc.y90_kernel(float(vox), grid, tissue, centre="reference")
            assert np.isnan(lit[c])
            m = np.ones(grid, bool)
            m[c] = False
            np.testing.assert_allclose(mine[m], ref[m], rtol=1e-14, atol=0)
            assert np.isfinite(mine[c])
        else:
            np.testing.assert_allclose(mine, ref, rtol=1e-14, atol=0)


def test_kernel_known_answers():
    k = json.load(open(os.path.join(GOLD, "kats.json")))
    y = orc.y90_kernel(1.0, (64, 64, 64), "water")
    assert y[33, 32, 32] == pytest.approx(k["y90_water_1mm_64_c+1"], rel=1e-14)
    assert y[32, 32, 32] == 1.0
    assert orc.y90_kernel(2.0, (64, 64, 64), "water")[33, 32, 32] == pytest.approx(k["y90_water_2mm_64_c+1"], rel=1e-14)
    assert orc.y90_kernel(1.0, (64, 64, 64), "lung")[33, 32, 32] == pytest.approx(k["y90_lung_1mm_64_c+1"], rel=1e-14)
    assert orc.y90_kernel(1.0, (64, 64, 64), "bone")[33, 32, 32] == pytest.approx(k["y90_bone_1mm_64_c+1"], rel=1e-14)
    l31 = orc.lu177_kernel(4.8, (31, 31, 31))
    assert l31[15, 15, 15] == pytest.approx(k["lu177_water_4.8mm_31_centre"], rel=1e-14)
    assert l31[16, 15, 15] == pytest.approx(k["lu177_water_4.8mm_31_c+1"], rel=1e-14)
    assert l31.sum() == pytest.approx(k["lu177_water_4.8mm_31_sum"], rel=1e-13)
    l81 = orc.lu177_kernel(1.0, (81, 81, 81))
    assert l81[41, 40, 40] == pytest.approx(k["lu177_water_1mm_81_c+1"], rel=1e-14)
    assert l81.sum() == pytest.approx(k["lu177_water_1mm_81_sum"], rel=1e-13)


def test_conv_golden_bit_exact_and_definition():
    z = np.load(os.path.join(GOLD, "conv_ref.npz"))
    names = sorted({k.split("|")[0] for k in z.files if k.endswith("|d")})
    assert len(names) == 5
    for name in names:
        a, k, d = z[name + "|a"], z[name + "|k"], z[name + "|d"]
        assert np.array_equal(orc.conv_reference(a, k), d), name            # same NumPy calls => same bits
        assert orc.rel_err_of_peak(orc.conv_bruteforce(a, k), d) < 1e-13     # FFT-free definition
        assert orc.rel_err_of_peak(orc.conv_reference_fast(a, k), d) < 1e-13
        assert d.sum() == pytest.approx(a.sum() * orc.crop_pad_kernel(k, a.shape).sum(), rel=1e-12)  # conservation


def test_time_integration_golden():
    z = np.load(os.path.join(GOLD, "conv_ref.npz"))
    maps, times, k, D, A = z["tp|maps"], z["tp|times"], z["tp|k"], z["tp|D"], z["tp|A"]
    assert np.array_equal(orc.absorbed_dose_trapezoid(list(maps), list(times), k), D)
    assert np.array_equal(orc.integrate_activity_trapezoid(list(maps), list(times)), A)
    w = orc.trapezoid_weights(times, 3600.0)
    one = orc.conv_reference(sum(wi * m for wi, m in zip(w, maps)), k)        # linearity: ONE convolution
    assert orc.rel_err_of_peak(one, D) < 1e-14
    w1 = orc.trapezoid_weights(times, 1.0)
    assert orc.rel_err_of_peak(sum(wi * m for wi, m in zip(w1, maps)), A) < 1e-15
    assert np.array_equal(orc.accumulated_activity_monoexp(z["a11|params"][0], z["a11|params"][1], 161.52), z["a11|acc"])
    assert np.array_equal(orc.accumulated_activity_monoexp(z["a11|params"][0], z["a11|params"][1], 161.52, 72.0), z["a11|acc72"])


def test_c1_example_known_answers():
    k = json.load(open(os.path.join(GOLD, "kats.json")))
    a = orc.sphere_activity()
    assert (a > 0).sum() == k["c1_sphere_voxels"] and a.sum() == k["c1_sum_a"]
    d = orc.conv_reference(a, orc.y90_kernel(1.0, (64, 64, 64), "water"))
    assert d.max() == pytest.approx(k["c1_max"], rel=1e-12)
    assert np.unravel_index(d.argmax(), d.shape) == (8, 8, 8)                # (24 + 32) mod 48
    assert d[0, 0, 0] == pytest.approx(k["c1_d000"], rel=1e-12)
    assert d[24, 24, 24] == pytest.approx(k["c1_d242424"], rel=1e-12)
    assert d.sum() == pytest.approx(k["c1_sum"], rel=1e-12)
    p = np.load(os.path.join(GOLD, "c1_ref.npz"))
    np.testing.assert_allclose(d[8], p["plane_x8"], rtol=1e-12)
    # the literal reference kernel (NaN centre) turns the whole map into NaN
    with np.errstate(invalid="ignore"):
        dn = orc.conv_reference(a, orc.y90_kernel(1.0, (64, 64, 64), "water", centre="reference"))
    assert np.isnan(dn).sum() == k["c1_literal_nan_voxels"]


def test_delta_response_is_rolled_cropped_kernel():
    rng = np.random.default_rng(3)
    for shape, kshape in (((12, 10, 9), (16, 4, 11)), ((8, 9, 10), (3, 5, 2))):
        k = rng.uniform(0, 1, kshape)
        a = np.zeros(shape)
        p = (5, 3, 7)
        a[p] = 1.0
        d = orc.conv_reference(a, k)
        np.testing.assert_allclose(d, np.roll(orc.crop_pad_kernel(k, shape), p, axis=(0, 1, 2)), atol=1e-14)


def test_same_mode_equals_scipy_for_odd_kernels():
    from scipy.signal import fftconvolve

    rng = np.random.default_rng(4)
    a, k = rng.uniform(0, 1, (9, 12, 7)), rng.uniform(0, 1, (5, 3, 7))
    np.testing.assert_allclose(orc.conv_same(a, k), fftconvolve(a, k, mode="same"), atol=1e-12)


@settings(max_examples=25, deadline=None)
@given(st.tuples(st.integers(1, 9), st.integers(1, 9), st.integers(1, 9)), st.tuples(st.integers(1, 6), st.integers(1, 6), st.integers(1, 6)),
       st.integers(0, 2**31 - 1))
def test_property_fft_equals_definition(shape, kshape, seed):
    rng = np.random.default_rng(seed)
    a, k = rng.normal(size=shape), rng.normal(size=kshape)
    ref = orc.conv_reference(a, k)
    assert np.max(np.abs(ref - orc.conv_bruteforce(a, k))) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    # linearity in the activity
    b = rng.normal(size=shape)
    np.testing.assert_allclose(orc.conv_reference(2 * a - 3 * b, k), 2 * ref - 3 * orc.conv_reference(b, k), atol=1e-10)


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_slab_decomposition_reference_logic(world):
    rng = np.random.default_rng(5)
    a, k = rng.uniform(0, 1, (13, 6, 7)), rng.uniform(0, 1, (4, 3, 9))
    np.testing.assert_allclose(orc.conv_reference_slabbed(a, k, world), orc.conv_reference(a, k), atol=1e-12)
    for ks in ((5, 3, 4), (4, 3, 5), (6, 2, 2)):  # odd and even kernel length along the slab axis
        k2 = rng.uniform(0, 1, ks)
        np.testing.assert_allclose(orc.conv_same_slabbed(a, k2, world), orc.conv_same(a, k2), atol=1e-12)


def test_density_and_hu():
    hu = np.array([-2000, -1000, -850, -700, -50, 0, 20, 40, 350, 1200, 5000], dtype=np.float64)
    rho = orc.hu_to_density(hu)
    assert rho[0] == rho[1] == 0.00129 and rho[-1] == 2.90
    assert rho[5] == 1.0 and rho[3] == 0.26 and rho[8] == 1.42
    assert rho[2] == pytest.approx((0.00129 + 0.26) / 2)
    d = orc.density_correct(np.ones(4), np.array([0.001, 0.05, 0.5, 2.0]), 1.0, 0.1, 0.01)
    np.testing.assert_allclose(d, [0.0, 10.0, 2.0, 0.5])
