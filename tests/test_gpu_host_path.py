"""GPU tests of the host-buffer (end-to-end) path: the staging engine behind the reference's calling convention
(pageable float64 / float32 ndarrays in, ndarray out - core/kernel_convolution.py:48-76), 16-bit stored activity
with rescale (io/dicom.py:27-47), int16 CT, and the device watchdog check on host-returning calls."""
import numpy as np
import pytest
import torch

from oracle import dose_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stager():
    from pyvoxeldosimetry_b200.engine import HostStager

    return HostStager.get(torch.device("cuda:0"))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16, np.uint16])
@pytest.mark.parametrize("n", [1, 4097, (1 << 20) + 3, 7 * (1 << 20) + 11, 50_000_001])
def test_stager_upload_is_exact(stager, dtype, n):
    rng = np.random.default_rng(n % 1000 + np.dtype(dtype).itemsize)
    if np.issubdtype(dtype, np.floating):
        a = (rng.standard_normal(n) * 1e3).astype(dtype)
    else:
        info = np.iinfo(dtype)
        a = rng.integers(info.min, info.max, size=n, endpoint=True, dtype=dtype)
    t = stager.upload(a)
    torch.cuda.synchronize()
    got = t.cpu().numpy()
    if dtype == np.uint16:
        got = got.view(np.uint16)
    want = a.astype(np.float32) if dtype == np.float64 else a
    assert got.dtype == want.dtype and np.array_equal(got, want)  # bit exact (float64 -> float32 is round-to-nearest on both sides)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [5, (1 << 20) + 3, 33_000_007])
def test_stager_download_is_exact(stager, dtype, n):
    t = torch.randn(n, device="cuda:0") * 1e3
    out = np.full(n, np.nan, dtype=dtype)
    stager.download(t, out)
    assert np.array_equal(out, t.cpu().numpy().astype(dtype))


def test_back_to_back_transfers_do_not_alias_ring_slots(stager):
    # many transfers in flight on different streams: every result must still be exact
    rng = np.random.default_rng(3)
    arrs = [rng.standard_normal(3_000_000 + 17 * i).astype(np.float64 if i % 2 else np.float32) for i in range(6)]
    outs = []
    streams = [torch.cuda.Stream("cuda:0") for _ in range(3)]
    for i, a in enumerate(arrs):
        with torch.cuda.stream(streams[i % 3]):
            outs.append(stager.upload(a))
    torch.cuda.synchronize()
    for a, t in zip(arrs, outs):
        assert np.array_equal(t.cpu().numpy(), a.astype(np.float32))
    back = np.empty(arrs[0].shape, np.float32)
    with torch.cuda.stream(streams[1]):
        stager.download(outs[0], back)
    assert np.array_equal(back, arrs[0].astype(np.float32))


def _calc(**cfg):
    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator

    return KernelConvolutionCalculator("Y90", "water", 1.0, config=dict({"kernel_grid": (15, 15, 15), "device": "cuda:0"}, **cfg))


def test_reference_calling_convention_float64_in_out():
    """float64 pageable ndarray in -> ndarray out through the staging engine, against the literal reference operator."""
    rng = np.random.default_rng(11)
    shape = (96, 80, 112)  # 6.9 MB as float64: above the staging threshold
    a = rng.uniform(0, 1e3, shape)
    a[40:50, 30:44, 50:70] = 2e6
    calc = _calc(output_dtype="float64")
    k = calc.kernel.astype(np.float32).astype(np.float64)
    ref = orc.conv_reference_fast(a.astype(np.float32).astype(np.float64), k)
    got = calc.calculate_dose_rate(activity_map=a, voxel_size=(1.0, 1.0, 1.0))
    assert got.dtype == np.float64 and got.shape == shape
    assert orc.rel_err_of_peak(got, ref) <= 1e-4
    got32 = _calc().calculate_dose_rate(a.astype(np.float32), (1.0, 1.0, 1.0))
    assert got32.dtype == np.float32 and orc.rel_err_of_peak(got32, ref) <= 1e-4
    # out= pageable ndarrays of either dtype are filled in place
    for dt in (np.float32, np.float64):
        out = np.full(shape, np.nan, dtype=dt)
        r = _calc().calculate_dose_rate(a, (1.0, 1.0, 1.0), out=out)
        assert r is out and orc.rel_err_of_peak(out, ref) <= 1e-4


def test_results_returned_from_pinned_cache_do_not_alias():
    """Two results held at once must be distinct buffers (the pinned blocks are recycled only after the ndarray dies)."""
    rng = np.random.default_rng(12)
    shape = (64, 64, 80)
    a1, a2 = rng.uniform(0, 1e3, shape).astype(np.float32), rng.uniform(0, 1e3, shape).astype(np.float32)
    calc = _calc()
    d1 = calc.calculate_dose_rate(a1, (1.0, 1.0, 1.0))
    keep = d1.copy()
    d2 = calc.calculate_dose_rate(a2, (1.0, 1.0, 1.0))
    assert not np.shares_memory(d1, d2)
    assert np.array_equal(d1, keep) and not np.array_equal(d1, d2)
    d1 += 1.0  # results are ordinary writable arrays


@pytest.mark.parametrize("dtype", [np.int16, np.uint16])
def test_stored_16bit_activity_with_rescale_and_int16_ct(dtype):
    rng = np.random.default_rng(13)
    shape = (80, 96, 128)
    stored = rng.integers(0, 30000, size=shape).astype(dtype)
    slope, intercept = 37.5, 12.0
    hu = rng.choice(np.array([-1000, -700, 32, 350], dtype=np.int16), size=shape)
    calc = _calc()
    got = calc.calculate_dose_rate(stored, (1.0, 1.0, 1.0), ct_hu=hu, rescale=(slope, intercept))
    a = np.float32(slope) * stored.astype(np.float32) + np.float32(intercept)
    k = calc.kernel.astype(np.float32).astype(np.float64)
    from pyvoxeldosimetry_b200.tissue.density import HU_KNOTS

    rho = orc.hu_to_density(hu.astype(np.float64), np.asarray(HU_KNOTS, dtype=np.float64))
    ref = orc.density_correct(orc.conv_reference_fast(a.astype(np.float64), k), rho, 1.0, 0.1, 0.0)
    assert orc.rel_err_of_peak(got, ref) <= 1e-4
    # the pipelined batch call takes the same stored volumes
    outs = calc.calculate_dose_rate_batch([stored, stored], (1.0, 1.0, 1.0), ct_hu=[hu, hu], rescale=(slope, intercept))
    for o in outs:
        assert orc.rel_err_of_peak(o, ref) <= 1e-4
    with pytest.raises(ValueError):
        calc.calculate_dose_rate(a, (1.0, 1.0, 1.0), rescale=(2.0, 0.0))


def test_user_kernel_survives_other_voxel_size_and_getter_is_read_only():
    calc = _calc()
    rng = np.random.default_rng(14)
    k = rng.uniform(0, 1, (7, 7, 7))
    calc.kernel = k
    a = rng.uniform(0, 1e3, (40, 40, 40)).astype(np.float32)
    ref = orc.conv_reference(a.astype(np.float64), k.astype(np.float32).astype(np.float64))
    got = calc.calculate_dose_rate(a, (2.0, 2.0, 3.0))  # voxel size != kernel_resolution: the user's kernel must still be used
    assert orc.rel_err_of_peak(got, ref) <= 1e-4
    with pytest.raises(ValueError):
        calc.kernel[3, 3, 3] = 0.0  # in-place edits would never reach the device
    k[3, 3, 3] = 123.0  # the caller's array is not aliased
    assert calc.kernel[3, 3, 3] != 123.0


@pytest.mark.parametrize("boundary", ["reference", "same"])
def test_host_pipeline_density_blocks_and_dose_blocks_overlap_correctly(boundary):
    """Volumes above 8 MB take the pipelined host call (density / CT uploaded plane block by plane block while finished
    dose blocks travel back): every input flavour must give the oracle's dose map."""
    from pyvoxeldosimetry_b200.tissue.density import HU_KNOTS

    rng = np.random.default_rng(15)
    shape = (150, 144, 120)  # 10.4 MB per float32 volume; 150 planes do not divide by the 8 blocks
    a = rng.uniform(0, 1e3, shape).astype(np.float32)
    a[60:90, 50:100, 40:80] = 2e6
    rho = rng.choice(np.array([0.00129, 0.26, 1.04, 1.42], dtype=np.float32), size=shape)
    hu = rng.choice(np.array([-1000, -700, 32, 350], dtype=np.int16), size=shape)
    calc = _calc(boundary=boundary)
    k = calc.kernel.astype(np.float32).astype(np.float64)
    conv = (orc.conv_reference_fast if boundary == "reference" else (lambda x, kk: orc.conv_same(x, kk, fast=True)))(a.astype(np.float64), k)
    ref_plain = conv
    ref_rho = orc.density_correct(conv, rho.astype(np.float64), 1.0, 0.1, 0.0)
    ref_hu = orc.density_correct(conv, orc.hu_to_density(hu.astype(np.float64), np.asarray(HU_KNOTS, dtype=np.float64)), 1.0, 0.1, 0.0)
    vs = (1.0, 1.0, 1.0)
    assert orc.rel_err_of_peak(calc.calculate_dose_rate(a, vs), ref_plain) <= 1e-4
    assert orc.rel_err_of_peak(calc.calculate_dose_rate(a.astype(np.float64), vs, tissue_densities=rho), ref_rho) <= 1e-4
    assert orc.rel_err_of_peak(calc.calculate_dose_rate(a, vs, tissue_densities=rho.astype(np.float64)), ref_rho) <= 1e-4
    assert orc.rel_err_of_peak(calc.calculate_dose_rate(a, vs, ct_hu=hu), ref_hu) <= 1e-4
    assert orc.rel_err_of_peak(calc.calculate_dose_rate(a, vs, ct_hu=hu.astype(np.float32)), ref_hu) <= 1e-4
    out = np.full(shape, np.nan, dtype=np.float64)
    assert calc.calculate_dose_rate(a, vs, tissue_densities=rho, out=out) is out and orc.rel_err_of_peak(out, ref_rho) <= 1e-4
    got64 = _calc(boundary=boundary, output_dtype="float64").calculate_dose_rate(a, vs, tissue_densities=rho)
    assert got64.dtype == np.float64 and orc.rel_err_of_peak(got64, ref_rho) <= 1e-4
    # multi-timepoint absorbed dose through the same pipeline
    maps = [a, (0.7 * a).astype(np.float32), (0.2 * a).astype(np.float32)]
    t = [2.0, 24.0, 72.0]
    w = orc.trapezoid_weights(t, 3600.0)
    acc = sum(wi * m.astype(np.float64) for wi, m in zip(w, maps))
    refD = (orc.conv_reference_fast if boundary == "reference" else (lambda x, kk: orc.conv_same(x, kk, fast=True)))(acc, k)
    gotD = calc.calculate_absorbed_dose(maps, t, vs, tissue_densities=rho)
    assert orc.rel_err_of_peak(gotD, orc.density_correct(refD, rho.astype(np.float64), 1.0, 0.1, 0.0)) <= 1e-4
    with pytest.raises(ValueError):
        calc.calculate_dose_rate(a, vs, tissue_densities=rho[:-1])
