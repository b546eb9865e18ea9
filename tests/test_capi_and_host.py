"""C-ABI surface and host-side logic.  No GPU compute here."""
import os
import re
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(REPO, "include", "pvdose.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pvd_[a-z0-9_]+)\s*\(", txt)))


def test_product_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    from pyvoxeldosimetry_b200._capi import SYMBOLS, PvdLib

    g.build()  # no-op when up to date; cross-compiles sm_100a without a GPU
    lib = PvdLib()  # libpvdose.so (CUDA build) must load without a GPU
    declared = _header_symbols()
    assert sorted(SYMBOLS) == declared, "binding and header disagree"
    for s in declared:
        assert hasattr(lib.dll, s), s
    assert lib.version() == 100
    out = subprocess.run(["nm", "-D", "--defined-only", lib.path], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (pvd_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_library_is_sm100a_and_has_no_cpu_fallback():
    so = os.path.join(REPO, "pyvoxeldosimetry_b200", "libpvdose.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    import torch

    if not torch.cuda.is_available():
        from pyvoxeldosimetry_b200 import DoseCalculator, KernelConvolutionCalculator

        with pytest.raises(RuntimeError, match="no CPU fallback"):
            KernelConvolutionCalculator("Y90", "water")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            DoseCalculator("Y90", "kernel", {"half_life": 64.1})


def test_missing_library_fails_loudly(tmp_path):
    from pyvoxeldosimetry_b200._capi import PvdLib, PvdoseLibraryError

    with pytest.raises(PvdoseLibraryError, match="no CPU fallback"):
        PvdLib(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle_or_emulator():
    pkg = os.path.join(REPO, "pyvoxeldosimetry_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f
                assert "emu_util" not in src and "libpvdose_emu" not in src, f


def test_trapezoid_weights_and_sampler_validation():
    from oracle import dose_oracle as orc
    from pyvoxeldosimetry_b200.core import ActivitySampler, trapezoid_weights

    t = [4.0, 24.0, 96.0, 168.0]
    np.testing.assert_allclose(trapezoid_weights(t, 3600.0), orc.trapezoid_weights(t, 3600.0), rtol=1e-15)
    assert trapezoid_weights([5.0]) == [0.0]
    assert trapezoid_weights([]) == []
    with pytest.raises(ValueError, match="Half-life must be positive"):  # activity_sampler.py:25-26
        ActivitySampler(0.0)
    with pytest.raises(ValueError, match="Invalid time units"):
        ActivitySampler(1.0, "days")
    s = ActivitySampler(64.1)
    w = s.dose_rate_weights([0.0, 10.0], integration_limit=None)
    assert w == [18000.0, 18000.0]
    w2 = s.dose_rate_weights([0.0, 10.0], integration_limit=1e9)
    assert w2[1] - w[1] == pytest.approx(64.1 / np.log(2) * 3600.0)
    ref = orc.integrate_dose_rates([np.ones(3), 2 * np.ones(3)], [0.0, 10.0], 50.0, 64.1)
    np.testing.assert_allclose(sum(wi * r for wi, r in zip(s.dose_rate_weights([0.0, 10.0], 50.0), [np.ones(3), 2 * np.ones(3)])), ref)


def test_generator_physics_matches_reference_golden():
    """radial_terms() of the host generators evaluated in NumPy == the real reference's kernels."""
    from pyvoxeldosimetry_b200.data.dose_kernels.generators import GENERATORS
    from oracle import dose_oracle as orc

    z = np.load(os.path.join(REPO, "tests", "golden", "kernels_ref.npz"))
    for key in z.files:
        nuc, tissue, vox, grid = key.split("|")
        grid = tuple(int(g) for g in grid.split("x"))
        beta, phot, sc = GENERATORS[nuc](tissue).radial_terms()
        r = orc.radial_grid(grid, (float(vox),) * 3)
        k = np.zeros(grid)
        for rng_, amp in beta:
            m = r <= rng_
            k[m] += amp * (1 - r[m] / rng_) ** 2 * np.exp(-2 * r[m] / rng_)
        m = r > 0
        for mu, amp in phot:
            k[m] += amp * np.exp(-mu * r[m] / 10) / (4 * np.pi * r[m] ** 2)
        k *= sc
        ref = z[key]
        ok = np.isfinite(ref)
        np.testing.assert_allclose(k[ok], ref[ok], rtol=1e-13)


def test_factory_errors_and_defaults():
    from pyvoxeldosimetry_b200.data.dose_kernels import KernelFactory

    f = KernelFactory()
    with pytest.raises(ValueError, match="Unsupported nuclide: Xx1"):
        f.get_kernel("Xx1", "water")
    with pytest.raises(ValueError, match="Supported:"):
        f.get_kernel("Tb161", "water")
    assert f._default_grid_sizes["Y90"] == (201, 201, 201) and f._default_grid_sizes["Lu177"] == (81, 81, 81)


def test_front_door_validation_without_gpu():
    from pyvoxeldosimetry_b200 import DoseCalculator

    with pytest.raises(ValueError, match="Unsupported calculation method"):
        DoseCalculator("Y90", "magic")
    with pytest.raises(NotImplementedError):
        DoseCalculator("Y90", "gate_monte_carlo")
    from pyvoxeldosimetry_b200.core.dosimetry_base import DosimetryCalculator

    class Dummy(DosimetryCalculator):
        def calculate_dose_rate(self, a, v):
            return a

        def calculate_absorbed_dose(self, a, t, v):
            return a[0]

    with pytest.raises(TypeError, match="Radionuclide must be a string"):  # dosimetry_base.py:30-31
        Dummy(90, None)
    d = Dummy("Y90", None, {"x": 1})
    cfg = d.get_config()
    cfg["x"] = 2
    assert d.config["x"] == 1


def test_shard_range_and_slab_geometry():
    from pyvoxeldosimetry_b200.multi_gpu import shard_range, slab_bounds, slab_geometry
    from emu_util import emu_lib

    assert [list(shard_range(10, 4, r)) for r in range(4)] == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert list(shard_range(64, 8, 3)) == list(range(24, 32))
    assert slab_bounds(1024, 8)[7] == (896, 1024)
    g = slab_geometry((1024, 1024, 800), (51, 51, 51), "same", 8, 3, emu_lib())
    assert (g["lo"], g["hi"], g["need_lo"], g["need_hi"]) == (384, 512, 359, 537)
    assert g["ex"]["out_lo"][0] == 50 and g["ex"]["out_n"] == (128, 1024, 800) and g["ex"]["m"][0] >= 178
    g = slab_geometry((1024, 1024, 800), (51, 51, 51), "reference", 8, 0, emu_lib())
    assert (g["need_lo"], g["need_hi"]) == (-50, 128) and g["ex"]["out_lo"] == (50, 0, 0) and g["ex"]["m"][1:] == (1024, 800)


def test_streaming_host_copies(tmp_path):
    """The staging threads' non-temporal copies / conversions (csrc/host_copy.h, plain C++): every head / tail alignment,
    float64 -> float32 rounds like NumPy's astype, float32 -> float64 is exact."""
    import ctypes

    src = tmp_path / "hc.cpp"
    src.write_text('#include "host_copy.h"\n'
                   'extern "C" void hc_copy(void* d, const void* s, size_t n) { pvd::stream_copy(d, s, n); }\n'
                   'extern "C" void hc_narrow(float* d, const double* s, size_t n) { pvd::stream_narrow(d, s, n); }\n'
                   'extern "C" void hc_widen(double* d, const float* s, size_t n) { pvd::stream_widen(d, s, n); }\n')
    so = tmp_path / "hc.so"
    csrc = os.path.join(REPO, "pyvoxeldosimetry_b200", "csrc")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", f"-I{csrc}", "-o", str(so), str(src)], check=True)
    dll = ctypes.CDLL(str(so))
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, 5000, dtype=np.uint8)
    for off_s in (0, 1, 7):
        for off_d in (0, 3, 4, 15):
            for n in (0, 1, 15, 16, 63, 64, 65, 200, 4096 + 13):
                dst = np.zeros(n + 64, dtype=np.uint8)
                dll.hc_copy(ctypes.c_void_p(dst.ctypes.data + off_d), ctypes.c_void_p(raw.ctypes.data + off_s), ctypes.c_size_t(n))
                assert np.array_equal(dst[off_d : off_d + n], raw[off_s : off_s + n])
                assert not dst[:off_d].any() and not dst[off_d + n :].any()
    d64 = np.concatenate([rng.standard_normal(1003) * 10.0 ** rng.integers(-30, 30, 1003), [0.0, -0.0, 1e300, -1e300, 1e-320, np.inf]])
    with np.errstate(over="ignore"):
        want32 = d64.astype(np.float32)
    for off in (0, 1, 2, 3):
        for n in (0, 1, 3, 4, 5, 8, 1009):
            buf = np.zeros(n + 8, dtype=np.float32)
            dll.hc_narrow(ctypes.c_void_p(buf.ctypes.data + 4 * off), ctypes.c_void_p(d64.ctypes.data), ctypes.c_size_t(n))
            assert np.array_equal(buf[off : off + n], want32[:n]) and not buf[:off].any() and not buf[off + n :].any()
            wide = np.zeros(n + 4, dtype=np.float64)
            dll.hc_widen(ctypes.c_void_p(wide.ctypes.data + 8 * (off & 1)), ctypes.c_void_p(want32.ctypes.data), ctypes.c_size_t(n))
            o = off & 1
            assert np.array_equal(wide[o : o + n], want32[:n].astype(np.float64)) and not wide[o + n :].any()
