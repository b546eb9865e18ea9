"""SURVEY.md section 8f rows - the steps either side of the convolution: per-voxel mono-exponential fit +
integral, CT artifact handling / tissue classes / density, dose-volume histogram, `.dat` kernels and NIfTI dose
maps.  CPU part: the oracle against the vectors the REAL reference produced (tests/golden/next_ref.npz,
oracle/gen_golden.py), the real kernels through the SIMT emulator, the host file formats.  GPU part: the
product API (CUDA through the C ABI) against the golden vectors and the oracle on larger seeded inputs.

Tolerances.  Fit: the reference is scipy's MINPACK Levenberg-Marquardt at its default ftol = xtol = 1.49e-8,
the device solves the same least-squares problem in float32; accumulated activity must agree to 1e-4 of its
peak (north-star tolerance), lambda / A0 of well-posed voxels to 1e-4 relative.  CT fill: float32 9^3 gather
vs scipy's separable float64 passes, 1e-3 HU absolute.  Tissue classes, DVH counts, `.dat` payload: exact.
"""
import os
import struct

import numpy as np
import pytest

from oracle import dose_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FIT_CASES = ("fit4", "fit3w", "fit6")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "next_ref.npz"))


def _check_fit(gold, name, params, acc):
    pr, ar = gold[name + "|params"], gold[name + "|acc"]
    assert np.isfinite(acc).all()
    assert np.max(np.abs(acc - ar)) <= 1e-4 * np.max(np.abs(ar))
    good = np.abs(pr[0]) > 1e-3 * np.abs(pr[0]).max()  # A0 ~ 0 leaves lambda undetermined
    np.testing.assert_allclose(params[1][good], pr[1][good], rtol=1e-4)
    np.testing.assert_allclose(params[0][good], pr[0][good], rtol=1e-4)
    zero = gold[name + "|maps"][:, 0, 0, 1]
    assert not zero.any() and params[0][0, 0, 1] == 0 and acc[0, 0, 1] == 0  # all-zero voxel -> A0 = 0, like the reference


# ----------------------------------------------------------------------------------------- oracle vs reference vectors
def test_oracle_fit_matches_reference_vectors(gold):
    for name in FIT_CASES:
        w = gold[name + "|weights"] if name + "|weights" in gold.files else None
        p, a = orc.fit_monoexp_curvefit(gold[name + "|times"], list(gold[name + "|maps"]), 161.52, w)
        np.testing.assert_allclose(p, gold[name + "|params"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(a, gold[name + "|acc"], rtol=1e-9, atol=1e-9)


def test_oracle_ct_dvh_dat_match_reference_vectors(gold, tmp_path):
    ct = gold["ct|hu"]
    np.testing.assert_array_equal(orc.handle_artifacts(ct), gold["ct|corrected"])
    for handle, tag in ((True, "comp"), (False, "comp_raw")):
        comp = orc.tissue_composition(ct, handle)
        for k, v in comp.items():
            np.testing.assert_array_equal(v.astype(np.uint8), gold[f"ct|{tag}|{k}"])
    dose, mask = gold["dvh|dose"], gold["dvh|mask"]
    for bins in (1000, 17):
        e, c = orc.calculate_dvh(dose, mask, bins)
        np.testing.assert_array_equal(e, gold[f"dvh|edges{bins}"])
        np.testing.assert_array_equal(c, gold[f"dvh|cum{bins}"])
    p = tmp_path / "k.dat"
    p.write_bytes(gold["dat|blob"].tobytes())
    k, md = orc.load_kernel_dat(p)
    np.testing.assert_array_equal(k, gold["dat|kernel"])
    assert md["created_by"] == "devhliu" and md["creation_date"] == "2025-02-08 09:50:56"


# ----------------------------------------------------------------------------------------- host file formats (no GPU)
def _io():
    """The io modules are pure host code; load them without importing the CUDA-facing package __init__."""
    import importlib.util

    mods = {}
    for name in ("kernel_dat", "nifti"):
        path = os.path.join(os.path.dirname(GOLD), "..", "pyvoxeldosimetry_b200", "io", name + ".py")
        spec = importlib.util.spec_from_file_location("pvd_io_" + name, os.path.abspath(path))
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    return mods["kernel_dat"], mods["nifti"]


def test_dat_kernel_reader_writer(gold, tmp_path):
    kd, _ = _io()
    p = tmp_path / "ref.dat"
    p.write_bytes(gold["dat|blob"].tobytes())
    k, md = kd.load_kernel(p)
    np.testing.assert_array_equal(k, gold["dat|kernel"])  # bytes laid out per core/utils.py:31-41, parsed by the real reader
    assert md["created_by"] == "devhliu" and md["creation_date"] == "2025-02-08 09:50:56"
    assert md["dimensions"].tolist() == [4, 5, 6] and md["voxel_size"] == np.float32(1.5) and md["scaling_factor"] == np.float32(2.0)
    # our writer -> the reference's reader (oracle restatement) and ours
    rng = np.random.default_rng(5)
    kk = rng.uniform(0, 3, (3, 7, 5)).astype(np.float32)
    q = tmp_path / "mine.dat"
    from datetime import datetime
    kd.save_kernel(q, kk, 2.5, total_energy=0.9337, scaling=4.0, created=datetime(2026, 10, 17, 1, 2, 3), user="builder")
    k1, m1 = orc.load_kernel_dat(q)
    k2, m2 = kd.load_kernel(q)
    np.testing.assert_array_equal(k1, k2)
    np.testing.assert_allclose(k1, kk, rtol=2e-7)
    assert m1["creation_date"] == m2["creation_date"] == "2026-10-17 01:02:03" and m2["created_by"] == "builder"
    assert os.path.getsize(q) == 80 + kk.size * 4
    # truncated / inconsistent files fail loudly
    (tmp_path / "short.dat").write_bytes(b"\x00" * 40)
    with pytest.raises(ValueError):
        kd.load_kernel(tmp_path / "short.dat")
    (tmp_path / "bad.dat").write_bytes(q.read_bytes()[:-8])
    with pytest.raises(ValueError):
        kd.load_kernel(tmp_path / "bad.dat")


def test_nifti_dose_map_layout_and_roundtrip(tmp_path):
    _, nf = _io()
    rng = np.random.default_rng(9)
    dose = rng.uniform(0, 50, (5, 6, 7))
    meta = {"radionuclide": "Y90", "method": "kernel", "arr": np.arange(3)}
    path = nf.save_dose_map(tmp_path / "dose", dose, (1.0, 2.0, 4.8), meta)
    assert path.endswith("dose.nii.gz")                      # suffix rule of core/utils.py:69-70
    assert meta["created_by"] == "devhliu" and meta["dimensions"] == (5, 6, 7)  # caller's dict is updated (:80-86)
    import gzip
    raw = gzip.open(path, "rb").read()
    # NIfTI-1 single-file layout
    assert struct.unpack_from("<i", raw, 0)[0] == 348 and raw[344:348] == b"n+1\x00"
    assert struct.unpack_from("<8h", raw, 40) == (3, 5, 6, 7, 1, 1, 1, 1)
    assert struct.unpack_from("<2h", raw, 70) == (16, 32)    # datatype float32, bitpix 32
    np.testing.assert_allclose(struct.unpack_from("<3f", raw, 80), (1.0, 2.0, 4.8), rtol=1e-7)
    vox = int(struct.unpack_from("<f", raw, 108)[0])
    assert vox % 16 == 0 and raw[348] == 1                   # extension flag set, data 16-byte aligned
    esize, ecode = struct.unpack_from("<2i", raw, 352)
    assert ecode == 44 and esize % 16 == 0 and 352 + esize == vox
    assert len(raw) == vox + dose.size * 4
    data = np.frombuffer(raw, "<f4", offset=vox).reshape(dose.shape, order="F")
    np.testing.assert_array_equal(data, dose.astype(np.float32))
    affine = np.array([struct.unpack_from("<4f", raw, o) for o in (280, 296, 312)])
    np.testing.assert_allclose(affine[:, :3], np.diag([1.0, 2.0, 4.8]), rtol=1e-7)
    np.testing.assert_allclose(affine[:, 3], -np.array([5, 6, 7]) * np.array([1.0, 2.0, 4.8]) / 2, rtol=1e-7)  # centred (:75-77)
    d2, m2 = nf.load_dose_map(path)
    np.testing.assert_array_equal(d2, dose.astype(np.float32))
    assert m2["radionuclide"] == "Y90" and m2["arr"] == [0, 1, 2] and m2["data_type"] == "float32"
    assert "Created by devhliu" in m2["description"] and m2["voxel_size"] == [1.0, 2.0, 4.8]
    # plain .nii with an explicit affine
    aff = np.eye(4)
    aff[:3, 3] = (10, 20, 30)
    p2 = nf.save_dose_map(str(tmp_path / "d.nii"), dose, (1, 1, 1), {}, affine=aff)
    d3, m3 = nf.load_dose_map(p2)
    assert p2.endswith("d.nii") and np.array_equal(d3, dose.astype(np.float32)) and m3["affine_matrix"][0][3] == 10
    with pytest.raises(ValueError):
        (tmp_path / "junk.nii").write_bytes(b"\x01" * 400)
        nf.load_dose_map(str(tmp_path / "junk.nii"))


# ----------------------------------------------------------------------------------------- real kernels, SIMT emulator
@pytest.fixture(scope="module")
def emu():
    from emu_util import emu_lib

    return emu_lib()


def test_emulated_fit_kernel_vs_reference_vectors(emu, gold):
    for name in FIT_CASES:
        maps = [np.ascontiguousarray(m, dtype=np.float32) for m in gold[name + "|maps"]]
        w = gold[name + "|weights"].tolist() if name + "|weights" in gold.files else None
        shape = maps[0].shape
        params = np.empty((2,) + shape, np.float32)
        acc = np.empty(shape, np.float32)
        emu.monoexp_fit([m.ctypes.data for m in maps], gold[name + "|times"].tolist(), w, np.log(2) / 161.52, 100 * 161.52,
                        params[0].ctypes.data, params[1].ctypes.data, acc.ctypes.data, acc.size)
        _check_fit(gold, name, params.astype(np.float64), acc.astype(np.float64))
    from pyvoxeldosimetry_b200._capi import PvdoseError
    with pytest.raises(PvdoseError):  # one time point cannot determine two parameters
        emu.monoexp_fit([maps[0].ctypes.data], [1.0], None, 0.1, 10.0, None, None, acc.ctypes.data, acc.size)


def _ranges():
    return [orc.TISSUE_HU_RANGES[k] for k in ("air", "lung", "soft_tissue", "bone", "water")]


def _check_ct(gold, corrected, rho, labels, handle=True):
    ref = gold["ct|corrected"] if handle else gold["ct|hu"]
    assert np.max(np.abs(corrected - ref)) <= 1e-3
    metal = gold["ct|hu"] > 2000
    if handle:
        assert metal.sum() >= 29 and np.array_equal(corrected[~metal], gold["ct|hu"].astype(np.float32)[~metal])
    tag = "comp" if handle else "comp_raw"
    for bit, k in enumerate(("air", "lung", "soft_tissue", "bone", "water")):
        got = (labels >> bit) & 1
        refm = gold[f"ct|{tag}|{k}"]
        # a voxel whose (filled) HU lies within float32 round-off of a range edge may legitimately flip
        edge = np.zeros(ref.shape, bool)
        for e in orc.TISSUE_HU_RANGES[k]:
            edge |= np.abs(ref - e) <= 1e-3
        assert np.array_equal(got[~edge], refm[~edge]), k
    np.testing.assert_allclose(rho, orc.hu_to_density(corrected), rtol=2e-6)


def test_emulated_ct_prepare_vs_reference_vectors(emu, gold):
    hu = np.ascontiguousarray(gold["ct|hu"], dtype=np.float32)
    for handle in (True, False):
        corrected, rho, labels = np.empty_like(hu), np.empty_like(hu), np.empty(hu.shape, np.uint8)
        emu.ct_prepare(hu.ctypes.data, hu.shape, 2000.0 if handle else float("inf"), orc.HU_KNOTS.tolist(), _ranges(),
                       corrected.ctypes.data, rho.ctypes.data, labels.ctypes.data)
        _check_ct(gold, corrected, rho, labels, handle)
    # a volume that is only 4-byte aligned takes the scalar load/store path
    buf = np.zeros(hu.size + 1, np.float32)
    hu_off = buf[1:].reshape(hu.shape)
    hu_off[...] = hu
    assert hu_off.ctypes.data % 16 == 4 or hu_off.ctypes.data % 16 != 0
    c2 = np.empty_like(hu)
    l2 = np.empty(hu.shape, np.uint8)
    emu.ct_prepare(hu_off.ctypes.data, hu.shape, 2000.0, None, _ranges(), c2.ctypes.data, None, l2.ctypes.data)
    emu.ct_prepare(hu.ctypes.data, hu.shape, 2000.0, None, _ranges(), corrected.ctypes.data, None, labels.ctypes.data)
    assert np.array_equal(c2, corrected) and np.array_equal(l2, labels)
    # more than 8 density segments -> the generic (loop) variant; must equal np.interp with clamped ends
    knots12 = [(-1000.0 + 300.0 * j, 0.001 + 0.21 * j + 0.01 * j * j) for j in range(12)]
    r12 = np.empty_like(hu)
    emu.ct_prepare(hu.ctypes.data, hu.shape, float("inf"), knots12, None, None, r12.ctypes.data, None)
    kx, ky = np.array(knots12).T
    np.testing.assert_allclose(r12, np.interp(hu.astype(np.float64), kx, ky), rtol=3e-6)
    # tiny volume: reflections longer than the axis
    t = np.full((2, 1, 3), 100.0, np.float32)
    t[0, 0, 1] = 2500.0
    out = np.empty_like(t)
    emu.ct_prepare(t.ctypes.data, t.shape, 2000.0, None, None, out.ctypes.data, None, None)
    np.testing.assert_allclose(out, orc.handle_artifacts(t), atol=1e-3)


def _ct_boundary_case(rng):
    """HU values on, one ulp below and one ulp above every range end and every density knot (+ NaN, +-inf, filler), with
    overlapping, touching, single-point, empty and half-infinite class ranges: what the merged interval table must get right."""
    knots = [(-1000.0, 0.00129), (-700.0, 0.26), (-100.0, 0.92), (0.0, 1.0), (40.0, 1.05), (350.0, 1.42), (1200.0, 1.92), (3000.0, 2.9)]
    ranges = [(-1000.0, -700.0), (-700.0, -100.0), (-150.0, 40.0), (40.0, 40.0), (41.0, 350.0), (300.0, float("inf")), (10.0, -10.0),
              (-float("inf"), -900.0)]
    pts = np.array([v for lo_hi in ranges for v in lo_hi if np.isfinite(v)] + [k[0] for k in knots], np.float32)
    near = np.concatenate([pts, np.nextafter(pts, np.float32(-np.inf)), np.nextafter(pts, np.float32(np.inf))])
    hu = np.concatenate([near, rng.uniform(-1500, 3500, 997).astype(np.float32),
                         np.array([np.nan, np.inf, -np.inf, 0.0, -0.0], np.float32)]).astype(np.float32)
    hu = np.concatenate([hu, np.zeros((-hu.size) % 4, np.float32)]).reshape(1, -1, 4)
    h64 = hu.astype(np.float64)
    lab = np.zeros(hu.shape, np.uint8)
    for c, (lo, hi) in enumerate(ranges):
        with np.errstate(invalid="ignore"):
            lab |= ((hu >= np.float32(lo)) & (hu <= np.float32(hi))).astype(np.uint8) << c
    kx, ky = np.array(knots).T
    rho = np.interp(h64, kx, ky)
    rho[np.isnan(h64)] = ky[0]  # a NaN voxel gets the first knot's density and no class (as the sum form did)
    return hu, knots, ranges, lab, rho


def test_emulated_ct_prepare_interval_table_boundaries(emu):
    hu, knots, ranges, lab, rho = _ct_boundary_case(np.random.default_rng(21))
    r, l = np.empty_like(hu), np.empty(hu.shape, np.uint8)
    emu.ct_prepare(hu.ctypes.data, hu.shape, float("inf"), knots, ranges, None, r.ctypes.data, l.ctypes.data)
    np.testing.assert_array_equal(l, lab)
    np.testing.assert_allclose(r, rho, rtol=3e-6, atol=1e-7)
    # > 31 cuts (12 knots + 8 ranges with distinct ends = up to 28; 24 knots + ranges -> the 64-entry table)
    knots24 = [(-1000.0 + 150.0 * j, 0.001 + 0.1 * j + 0.004 * j * j) for j in range(24)]
    emu.ct_prepare(hu.ctypes.data, hu.shape, float("inf"), knots24, ranges, None, r.ctypes.data, l.ctypes.data)
    kx, ky = np.array(knots24).T
    want = np.interp(hu.astype(np.float64), kx, ky)
    want[np.isnan(hu)] = ky[0]
    np.testing.assert_array_equal(l, lab)
    np.testing.assert_allclose(r, want, rtol=3e-6, atol=1e-7)


def _check_dvh(gold, fn):
    dose, mask = gold["dvh|dose"], gold["dvh|mask"]
    for bins in (1000, 17):
        e, c = fn(dose, mask, bins)
        np.testing.assert_array_equal(e, gold[f"dvh|edges{bins}"])
        np.testing.assert_array_equal(c, gold[f"dvh|cum{bins}"])  # bit-identical: same counts, same float64 cumsum


def _emu_dvh(emu):
    def fn(dose, mask, bins):
        dose = np.ascontiguousarray(dose, np.float32)
        m = np.ascontiguousarray(mask > 0).view(np.uint8)
        scratch = np.zeros(16, np.uint8)
        mn, mx, cnt = emu.roi_minmax(dose.ctypes.data, m.ctypes.data, False, dose.size, scratch.ctypes.data)
        assert (mn, mx, cnt) == (dose[mask > 0].min(), dose[mask > 0].max(), int((mask > 0).sum()))
        edges = np.histogram_bin_edges(np.array([mn, mx], np.float32), bins=bins)
        hist = np.empty(bins, np.uint64)
        emu.dvh_histogram(dose.ctypes.data, m.ctypes.data, False, dose.size, edges.ctypes.data, bins, float(edges[0]), float(edges[-1]),
                          hist.ctypes.data)
        np.testing.assert_array_equal(hist, np.histogram(dose[mask > 0], bins=bins)[0])
        return edges[1:], 1.0 - np.cumsum(hist) / cnt
    return fn


def test_emulated_dvh_vs_reference_vectors(emu, gold):
    _check_dvh(gold, _emu_dvh(emu))
    # negative doses, a float32 mask, a constant ROI (numpy widens the range by +-0.5) and > 4096 bins (global atomics)
    rng = np.random.default_rng(3)
    dose = rng.normal(0, 5, (6, 5, 4)).astype(np.float32)
    maskf = rng.uniform(-1, 1, dose.shape).astype(np.float32)
    scratch = np.zeros(16, np.uint8)
    mn, mx, cnt = emu.roi_minmax(dose.ctypes.data, maskf.ctypes.data, True, dose.size, scratch.ctypes.data)
    sel = dose[maskf > 0]
    assert (mn, mx, cnt) == (sel.min(), sel.max(), sel.size) and mn < 0
    for bins in (5000, 3):
        edges = np.histogram_bin_edges(np.array([mn, mx], np.float32), bins=bins)
        hist = np.empty(bins, np.uint64)
        emu.dvh_histogram(dose.ctypes.data, maskf.ctypes.data, True, dose.size, edges.ctypes.data, bins, float(edges[0]), float(edges[-1]),
                          hist.ctypes.data)
        np.testing.assert_array_equal(hist, np.histogram(sel, bins=bins)[0])
    # 4-byte aligned views: scalar path, same counts
    d1, m1 = dose.ravel()[1:], np.ascontiguousarray(maskf.ravel()[1:] > 0).view(np.uint8)
    mn, mx, cnt = emu.roi_minmax(d1.ctypes.data, m1.ctypes.data, False, d1.size, scratch.ctypes.data)
    sel1 = d1[m1 > 0]
    assert d1.ctypes.data % 16 != 0 and (mn, mx, cnt) == (sel1.min(), sel1.max(), sel1.size)
    edges = np.histogram_bin_edges(np.array([mn, mx], np.float32), bins=7)
    hist = np.empty(7, np.uint64)
    emu.dvh_histogram(d1.ctypes.data, m1.ctypes.data, False, d1.size, edges.ctypes.data, 7, float(edges[0]), float(edges[-1]), hist.ctypes.data)
    np.testing.assert_array_equal(hist, np.histogram(sel1, bins=7)[0])
    const = np.full(10, 7.0, np.float32)
    ones = np.ones(10, np.uint8)
    edges = np.histogram_bin_edges(const, bins=4)
    hist = np.empty(4, np.uint64)
    emu.dvh_histogram(const.ctypes.data, ones.ctypes.data, False, 10, edges.ctypes.data, 4, float(edges[0]), float(edges[-1]), hist.ctypes.data)
    np.testing.assert_array_equal(hist, np.histogram(const, bins=4)[0])
    empty = np.zeros(10, np.uint8)
    assert emu.roi_minmax(const.ctypes.data, empty.ctypes.data, False, 10, scratch.ctypes.data)[2] == 0


def _edge_hugging_doses(first, last, bins, rng):
    """Doses on, one ulp below and one ulp above every float32 bin edge, plus uniform filler: the voxels for which the
    histogram's reciprocal-multiply guess and numpy's edge comparisons could disagree."""
    edges = np.histogram_bin_edges(np.array([first, last], np.float32), bins=bins).astype(np.float32)
    pick = edges if bins <= 4096 else edges[rng.integers(0, bins + 1, 4096)]
    near = np.concatenate([pick, np.nextafter(pick, np.float32(-np.inf)), np.nextafter(pick, np.float32(np.inf))])
    filler = rng.uniform(first, last, 5000).astype(np.float32)
    d = np.concatenate([near, filler, np.array([first, last], np.float32)]).astype(np.float32)
    return np.clip(d, np.float32(first), np.float32(last))


DVH_EDGE_CASES = [(0.0, 50.0, 1000), (1000.0, 1001.0, 1000), (-3.0, 7.5, 37), (0.0, 1e-3, 4096), (100.0, 5000.0, 65536), (-1e6, 1e6, 999)]


def test_emulated_dvh_edge_hugging_doses_match_numpy(emu):
    rng = np.random.default_rng(12)
    for first, last, bins in DVH_EDGE_CASES:
        d = _edge_hugging_doses(first, last, bins, rng)[:6000]  # the emulator runs one host thread per CUDA thread
        m = np.ones(d.size, np.uint8)
        edges = np.histogram_bin_edges(np.array([d.min(), d.max()], np.float32), bins=bins)
        hist = np.empty(bins, np.uint64)
        emu.dvh_histogram(d.ctypes.data, m.ctypes.data, False, d.size, edges.ctypes.data, bins, float(edges[0]), float(edges[-1]), hist.ctypes.data)
        np.testing.assert_array_equal(hist, np.histogram(d, bins=bins)[0], err_msg=str((first, last, bins)))


@pytest.mark.gpu
def test_gpu_dvh_edge_hugging_doses_match_numpy():
    import torch
    from pyvoxeldosimetry_b200 import engine

    rng = np.random.default_rng(13)
    dev = torch.device("cuda:0")
    for first, last, bins in DVH_EDGE_CASES:
        d = _edge_hugging_doses(first, last, bins, rng)
        d = np.concatenate([d, rng.uniform(first, last, 5 << 20).astype(np.float32)]).astype(np.float32)  # > 2^22: lane-private kernel for <= 1536 bins
        d = np.clip(d, np.float32(first), np.float32(last))
        edges = np.histogram_bin_edges(np.array([d.min(), d.max()], np.float32), bins=bins)
        hist = engine.dvh_histogram(torch.from_numpy(d).to(dev), torch.ones(d.size, dtype=torch.uint8, device=dev),
                                    torch.from_numpy(edges.astype(np.float32)).to(dev))
        np.testing.assert_array_equal(hist.cpu().numpy().astype(np.int64), np.histogram(d, bins=bins)[0], err_msg=str((first, last, bins)))


# ----------------------------------------------------------------------------------------- GPU: product API
@pytest.mark.gpu
def test_gpu_fit_api_vs_reference_vectors_and_oracle(gold):
    import torch
    from pyvoxeldosimetry_b200 import TimeCurveFitting

    tcf = TimeCurveFitting(161.52)
    for name in FIT_CASES:
        w = gold[name + "|weights"].tolist() if name + "|weights" in gold.files else None
        params, acc = tcf.fit_time_activity_curve(gold[name + "|times"].tolist(), list(gold[name + "|maps"]), w)
        assert params.shape == gold[name + "|params"].shape and params.dtype == np.float32
        _check_fit(gold, name, params.astype(np.float64), acc.astype(np.float64))
    # larger seeded volume against the literal scipy loop (2400 voxels ~ 1 s of curve_fit), device tensors in/out
    rng = np.random.default_rng(21)
    shape, times = (20, 12, 10), [4.0, 24.0, 96.0, 168.0]
    lam0 = np.log(2) / 161.52
    A0, lam = rng.uniform(1e2, 1e6, shape), lam0 * rng.uniform(0.7, 4.0, shape)
    maps = [A0 * np.exp(-lam * t) * (1 + 0.05 * rng.standard_normal(shape)) for t in times]
    pr, ar = orc.fit_monoexp_curvefit(times, maps, 161.52)
    dev = [torch.from_numpy(m.astype(np.float32)).cuda() for m in maps]
    pd, ad = tcf.fit_time_activity_curve(times, dev)
    assert pd.is_cuda and ad.is_cuda
    assert np.max(np.abs(ad.cpu().numpy() - ar)) <= 1e-4 * np.abs(ar).max()
    np.testing.assert_allclose(pd[1].cpu().numpy(), pr[1], rtol=1e-4)
    # the fitted integral feeds the convolution: dose from fitted accumulated activity == oracle pipeline
    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator
    calc = KernelConvolutionCalculator("Lu177", "water", 4.8, config={"kernel_grid": (9, 9, 9)})
    dose = calc.calculate_dose_rate(ad, (4.8, 4.8, 4.8))
    dose = dose.cpu().numpy() if hasattr(dose, "cpu") else dose
    ref = orc.conv_reference(ar, calc.kernel.astype(np.float32).astype(np.float64))
    assert orc.rel_err_of_peak(dose, ref) <= 1e-4
    with pytest.raises(ValueError):
        tcf.fit_time_activity_curve([1.0], [maps[0]])
    with pytest.raises(ValueError):
        tcf.fit_time_activity_curve(times, maps[:3])


@pytest.mark.gpu
def test_gpu_tissue_composition_vs_reference_vectors(gold):
    import torch
    from pyvoxeldosimetry_b200.tissue import TissueComposition

    tc = TissueComposition()
    ct = gold["ct|hu"]
    for handle in (True, False):
        comp = tc.calculate_composition(ct, handle_artifacts=handle)
        assert list(comp) == ["air", "lung", "soft_tissue", "bone", "water"] and comp["air"].dtype == np.float64
        labels = sum((comp[k].astype(np.uint8) << b) for b, k in enumerate(comp))
        corrected = tc._handle_artifacts(ct) if handle else ct.astype(np.float32)
        rho = tc.density_map(ct, handle_artifacts=handle)
        _check_ct(gold, corrected, rho, labels, handle)
    # full-size CT slab, device in / device out, sparse implants: the fill must equal scipy's separable filter there
    rng = np.random.default_rng(4)
    big = rng.uniform(-1000, 1800, (96, 80, 64)).astype(np.float32)
    idx = rng.integers(0, big.size, 500)
    big.reshape(-1)[idx] = 3000.0
    out = tc._handle_artifacts(torch.from_numpy(big).cuda())
    assert out.is_cuda
    assert np.max(np.abs(out.cpu().numpy() - orc.handle_artifacts(big))) <= 1e-3


@pytest.mark.gpu
def test_gpu_dvh_vs_reference_vectors_and_numpy(gold):
    import torch
    from pyvoxeldosimetry_b200.core.utils import calculate_dvh

    _check_dvh(gold, calculate_dvh)
    # 256^3 dose map resident on the device: counts identical to numpy on the same float32 doses
    rng = np.random.default_rng(6)
    dose = (rng.gamma(2.0, 4.0, (256, 256, 256)) * 10).astype(np.float32)
    mask = np.zeros(dose.shape, np.uint8)
    mask[40:200, 30:220, 50:180] = 1
    e, c = calculate_dvh(torch.from_numpy(dose).cuda(), torch.from_numpy(mask).cuda(), 1000)
    er, cr = orc.calculate_dvh(dose, mask, 1000)
    np.testing.assert_array_equal(e, er)
    np.testing.assert_array_equal(c, cr)
    with pytest.raises(ValueError, match="empty"):
        calculate_dvh(dose[:4, :4, :4], np.zeros((4, 4, 4)), 10)
    with pytest.raises(ValueError, match="same dimensions"):
        calculate_dvh(dose[:4, :4, :4], np.zeros((4, 4, 5)), 10)


# ---- NIfTI-1 writer pinned against the standard's own struct layout (nibabel is absent in this image) ----
# nifti_1_header of nifti1.h (NIfTI-1.1, 348 bytes), field by field, as a packed little-endian NumPy record.
NIFTI1_STRUCT = np.dtype([
    ("sizeof_hdr", "<i4"), ("data_type", "S10"), ("db_name", "S18"), ("extents", "<i4"), ("session_error", "<i2"), ("regular", "S1"),
    ("dim_info", "u1"), ("dim", "<i2", (8,)), ("intent_p1", "<f4"), ("intent_p2", "<f4"), ("intent_p3", "<f4"), ("intent_code", "<i2"),
    ("datatype", "<i2"), ("bitpix", "<i2"), ("slice_start", "<i2"), ("pixdim", "<f4", (8,)), ("vox_offset", "<f4"), ("scl_slope", "<f4"),
    ("scl_inter", "<f4"), ("slice_end", "<i2"), ("slice_code", "u1"), ("xyzt_units", "u1"), ("cal_max", "<f4"), ("cal_min", "<f4"),
    ("slice_duration", "<f4"), ("toffset", "<f4"), ("glmax", "<i4"), ("glmin", "<i4"), ("descrip", "S80"), ("aux_file", "S24"),
    ("qform_code", "<i2"), ("sform_code", "<i2"), ("quatern_b", "<f4"), ("quatern_c", "<f4"), ("quatern_d", "<f4"), ("qoffset_x", "<f4"),
    ("qoffset_y", "<f4"), ("qoffset_z", "<f4"), ("srow_x", "<f4", (4,)), ("srow_y", "<f4", (4,)), ("srow_z", "<f4", (4,)),
    ("intent_name", "S16"), ("magic", "S4"),
])


def test_nifti_header_bytes_match_the_nifti1_struct(tmp_path):
    """Byte-for-byte: the 348 header bytes the writer emits == the record the NIfTI-1 standard defines, filled with the
    values nibabel's Nifti1Image(float32 data, affine) + header extension 44 would carry (reference core/utils.py:88-107)."""
    _, nf = _io()
    assert NIFTI1_STRUCT.itemsize == 348
    rng = np.random.default_rng(44)
    dose = rng.uniform(0, 5, (7, 5, 3)).astype(np.float32)
    vs = (2.0, 1.5, 3.0)
    meta = {"radionuclide": "Lu177"}
    path = nf.save_dose_map(str(tmp_path / "pin.nii"), dose, vs, meta)
    raw = open(path, "rb").read()
    got = np.frombuffer(raw[:348], dtype=NIFTI1_STRUCT)[0]
    want = np.zeros((), dtype=NIFTI1_STRUCT)
    want["sizeof_hdr"] = 348
    want["dim"] = (3, 7, 5, 3, 1, 1, 1, 1)
    want["datatype"], want["bitpix"] = 16, 32                    # DT_FLOAT32
    want["pixdim"] = (1.0, 2.0, 1.5, 3.0, 1.0, 1.0, 1.0, 1.0)     # qfac, voxel sizes
    want["scl_slope"] = want["scl_inter"] = np.nan                # "no scaling" on disk
    want["xyzt_units"] = 2                                        # NIFTI_UNITS_MM
    want["descrip"] = b"Created by devhliu at 2025-02-08 09:50:56"
    want["qform_code"], want["sform_code"] = 0, 2                 # NIFTI_XFORM_ALIGNED_ANAT
    aff = np.diag([2.0, 1.5, 3.0, 1.0])
    aff[:3, 3] = np.array(dose.shape) * np.array(vs) / -2.0       # reference default affine, core/utils.py:72-77
    want["srow_x"], want["srow_y"], want["srow_z"] = aff[0], aff[1], aff[2]
    want["magic"] = b"n+1"
    content = open(path, "rb").read()
    esize = int(np.frombuffer(content[352:356], "<i4")[0])
    assert np.frombuffer(content[356:360], "<i4")[0] == 44 and esize % 16 == 0 and content[348:352] == b"\x01\x00\x00\x00"
    want["vox_offset"] = 352 + esize
    assert got.tobytes() == want.tobytes(), [n for n in NIFTI1_STRUCT.names if got[n].tobytes() != want[n].tobytes()]
    # voxel data: Fortran order float32 right at vox_offset
    data = np.frombuffer(content, "<f4", count=dose.size, offset=352 + esize).reshape(dose.shape, order="F")
    assert np.array_equal(data, dose)


@pytest.mark.gpu
def test_gpu_ct_prepare_interval_table_boundaries():
    import torch
    from pyvoxeldosimetry_b200 import engine

    hu, knots, ranges, lab, rho = _ct_boundary_case(np.random.default_rng(22))
    big = np.tile(hu, (64, 1, 1))  # enough groups for several warps per block
    _, r, l = engine.ct_prepare(torch.from_numpy(big).cuda(), float("inf"), knots, ranges, want_corrected=False)
    np.testing.assert_array_equal(l.cpu().numpy(), np.tile(lab, (64, 1, 1)))
    np.testing.assert_allclose(r.cpu().numpy(), np.tile(rho, (64, 1, 1)), rtol=3e-6, atol=1e-7)
