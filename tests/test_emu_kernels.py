"""The REAL kernel sources (compiled for the CPU SIMT emulator, see tests/emu_util.py) driven through the
REAL C ABI against the oracle: index math, radix schedules, crop/pad, fused epilogues, error codes.
No GPU needed; the numerics bar is the same 1e-4 of peak (observed ~1e-7)."""
import numpy as np
import pytest

from emu_util import EmuConv, emu_lib
from oracle import dose_oracle as orc

TOL = 1e-4


@pytest.fixture(scope="module")
def lib():
    return emu_lib()


def _rand(shape, kshape, seed):
    rng = np.random.default_rng(seed)
    a = rng.uniform(0, 1e3, shape)
    a[tuple(s // 2 for s in shape)] = 2e6
    return a, rng.uniform(0, 1, kshape)


CASES = [
    ((16, 12, 20), (5, 7, 3)), ((10, 9, 8), (12, 4, 11)), ((7, 11, 13), (3, 3, 3)), ((8, 8, 8), (1, 1, 1)),
    ((12, 10, 14), (4, 6, 2)), ((1, 1, 1), (1, 1, 1)), ((2, 3, 1), (3, 2, 2)), ((25, 9, 27), (3, 3, 3)),
    ((6, 49, 30), (2, 2, 2)), ((32, 17, 64), (5, 5, 5)), ((5, 34, 38), (3, 3, 3)),
]


@pytest.mark.parametrize("shape,kshape", CASES)
@pytest.mark.parametrize("boundary", [0, 1])
def test_conv_vs_oracle(lib, shape, kshape, boundary):
    a, k = _rand(shape, kshape, hash((shape, kshape)) % 2**32)
    p = EmuConv(lib, shape, kshape, boundary)
    p.set_kernel(k)
    got = p.execute([a])
    p.close()
    a32, k32 = a.astype(np.float32).astype(np.float64), k.astype(np.float32).astype(np.float64)
    ref = orc.conv_reference(a32, k32) if boundary == 0 else orc.conv_same(a32, k32)
    assert orc.rel_err_of_peak(got, ref) <= TOL


def test_time_weights_density_scale(lib):
    rng = np.random.default_rng(11)
    shape, kshape = (10, 12, 9), (3, 5, 3)
    maps = [rng.uniform(0, 1e3, shape) for _ in range(4)]
    times = [4.0, 24.0, 96.0, 168.0]
    k = rng.uniform(0, 1, kshape)
    rho = rng.choice([0.00129, 0.26, 1.04, 1.42], size=shape)
    w = orc.trapezoid_weights(times, 3600.0)
    p = EmuConv(lib, shape, kshape, 1)
    p.set_kernel(k)
    got = p.execute(maps, [float(x) for x in w], rho, 1.0, 0.1, 0.01, 0.5)
    p.close()
    f = lambda x: np.asarray(x, np.float32).astype(np.float64)
    acc = sum(np.float64(np.float32(wi)) * f(m) for wi, m in zip(w, maps))
    ref = orc.density_correct(0.5 * orc.conv_same(acc, f(k)), f(rho), 1.0, 0.1, 0.01)
    assert orc.rel_err_of_peak(got, ref) <= TOL
    assert np.all(got[rho < 0.01] == 0.0)


def test_golden_reference_vectors(lib):
    import os

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "conv_ref.npz"))
    for name in sorted({k.split("|")[0] for k in z.files if k.endswith("|d")}):
        a, k, d = z[name + "|a"], z[name + "|k"], z[name + "|d"]
        p = EmuConv(lib, a.shape, k.shape, 0)
        p.set_kernel(k)
        got = p.execute([a])
        p.close()
        assert orc.rel_err_of_peak(got, d) <= TOL, name


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("boundary", ["reference", "same"])
def test_slab_plans(lib, world, boundary):
    """Expert plans (slab + halo in, interior out) stitched over `world` slabs == whole-volume result."""
    from pyvoxeldosimetry_b200.multi_gpu import slab_geometry

    rng = np.random.default_rng(12)
    shape, kshape = (13, 6, 10), (4, 3, 5)
    a, k = rng.uniform(0, 1e3, shape), rng.uniform(0, 1, kshape)
    f = lambda x: np.asarray(x, np.float32).astype(np.float64)
    ref = orc.conv_reference(f(a), f(k)) if boundary == "reference" else orc.conv_same(f(a), f(k))
    out = np.empty(shape, np.float32)
    for r in range(world):
        g = slab_geometry(shape, kshape, boundary, world, r, lib)
        idx = np.arange(g["need_lo"], g["need_hi"])
        if boundary == "reference":
            local = a[idx % shape[0]]
        else:
            local = np.zeros((len(idx),) + shape[1:])
            ok = (idx >= 0) & (idx < shape[0])
            local[ok] = a[idx[ok]]
        kc = g["kcrop"]
        p = EmuConv(lib, g["n"], kc, ex=g["ex"])
        p.set_kernel(k[: kc[0], : kc[1], : kc[2]])
        out[g["lo"] : g["hi"]] = p.execute([local])
        p.close()
    assert orc.rel_err_of_peak(out, ref) <= TOL


def test_error_codes(lib):
    from pyvoxeldosimetry_b200._capi import PvdoseError

    with pytest.raises(PvdoseError) as e:
        lib.plan_create((0, 4, 4), (3, 3, 3), 0)
    assert e.value.code == -1
    with pytest.raises(PvdoseError) as e:
        lib.plan_create((4, 4, 4), (3, 3, 3), 7)
    assert e.value.code == -1
    plan = lib.plan_create((4, 4, 4), (3, 3, 3), 0)
    x = np.zeros((4, 4, 4), np.float32)
    with pytest.raises(PvdoseError) as e:  # workspace not set
        lib.conv_execute(plan, [x.ctypes.data], None, None, 1, 0.1, 0, 1, x.ctypes.data)
    assert e.value.code == -3
    lib.plan_destroy(plan)
    p = EmuConv(lib, (4, 4, 4), (3, 3, 3), 0)
    with pytest.raises(PvdoseError) as e:  # kernel not set
        p.execute([x])
    assert e.value.code == -3
    k = np.ones((3, 3, 3), np.float32)
    k[1, 1, 1] = np.nan
    with pytest.raises(PvdoseError) as e:  # the reference's own Y90 kernel is non-finite: reject loudly
        p.set_kernel(k)
    assert e.value.code == -4
    p.set_kernel(np.ones((3, 3, 3)))
    with pytest.raises(PvdoseError) as e:
        lib.conv_execute(p.plan, [x.ctypes.data] * 17, None, None, 1, 0.1, 0, 1, x.ctypes.data)
    assert e.value.code == -1
    p.close()


def test_good_fft_size(lib):
    for n in (1, 2, 17, 48, 100, 400, 425, 537, 1000):
        m = lib.good_fft_size(n)
        assert m >= n
        r = m
        for q in (2, 3, 5, 7):
            while r % q == 0:
                r //= q
        assert r == 1


def test_elementwise_ops(lib):
    rng = np.random.default_rng(13)
    # radial model == oracle generators (finite centre)
    from pyvoxeldosimetry_b200.data.dose_kernels.generators import GENERATORS

    for nuc, vox, grid, tissue in (("Y90", 1.0, (9, 8, 7), "bone"), ("Lu177", 4.8, (7, 7, 7), "lung"), ("Y90", (1.0, 2.0, 0.5), (6, 5, 9), "water")):
        gen = GENERATORS[nuc](tissue)
        beta, phot, sc = gen.radial_terms()
        out = np.empty(grid, np.float32)
        sp = (vox,) * 3 if np.isscalar(vox) else vox
        lib.kernel_eval_radial(beta, phot, sc, sp, grid, out.ctypes.data)
        ref = orc.make_kernel(nuc, vox, grid, tissue)
        np.testing.assert_allclose(out, ref, rtol=2e-7, atol=0)
    hu = rng.uniform(-1200, 3500, 1000).astype(np.float32)
    rho = np.empty_like(hu)
    lib.hu_to_density(hu.ctypes.data, False, orc.HU_KNOTS.tolist(), rho.ctypes.data, hu.size)
    np.testing.assert_allclose(rho, orc.hu_to_density(hu), rtol=2e-6)
    hi = rng.integers(-1100, 3200, 1000).astype(np.int16)
    lib.hu_to_density(hi.ctypes.data, True, orc.HU_KNOTS.tolist(), rho.ctypes.data, hi.size)
    np.testing.assert_allclose(rho, orc.hu_to_density(hi), rtol=2e-6)
    vols = [rng.uniform(0, 1, 777).astype(np.float32) for _ in range(5)]
    w = [0.5, 1.0, 2.0, -1.0, 3.0]
    out = np.empty(777, np.float32)
    lib.weighted_sum([v.ctypes.data for v in vols], w, out.ctypes.data, 777)
    np.testing.assert_allclose(out, sum(wi * v.astype(np.float64) for wi, v in zip(w, vols)), rtol=1e-6)
    a0, lam = rng.uniform(0, 1e4, 500).astype(np.float32), rng.uniform(1e-3, 1e-1, 500).astype(np.float32)
    lib.monoexp_integral(a0.ctypes.data, lam.ctypes.data, 72.0, out.ctypes.data, 500)
    np.testing.assert_allclose(out[:500], orc.accumulated_activity_monoexp(a0, lam, 161.52, 72.0), rtol=2e-6)
    d, r = rng.uniform(0, 1, 300).astype(np.float32), rng.uniform(0, 2, 300).astype(np.float32)
    lib.density_scale(d.ctypes.data, r.ctypes.data, 1.0, 0.1, 0.05, 2.0, out.ctypes.data, 300)
    np.testing.assert_allclose(out[:300], orc.density_correct(2.0 * d.astype(np.float64), r, 1.0, 0.1, 0.05), rtol=1e-6)


# ---- size-specialised + persistent pipelined kernels (lengths on the menu: 512, 256, 400) ----------
FAST_CASES = [
    ((512, 3, 6), (5, 2, 3), 0, 2, True),      # x pass: cols_pipe_kernel<512> CONV, ntz == 1
    ((5, 40, 400), (2, 5, 7), 0, 1, True),     # rows pipe kernels <400>, 7 tiles over 3 emulated SMs
    ((3, 256, 5), (3, 9, 3), 0, 2, False),     # y pass <256>
    ((256, 2, 256), (4, 2, 9), 0, 1, True),    # x <256>, rows <256>, several tiles per CTA
    ((4, 33, 376), (1, 3, 41), 1, 1, True),    # same mode: 376 -> 400 partial cp.async zero fill
    ((3, 35, 379), (1, 3, 41), 1, 1, True),    # n2 % 4 != 0 -> unaligned rows fall back to the LDG kernel
    ((500, 2, 4), (25, 1, 3), 1, 1, False),    # same mode along x: 500 -> 512, crop offset 12
    ((1152, 2, 4), (5, 1, 2), 0, 1, True),     # slab-decomposition sizes: x <1152> 8*12*12 (one tile per CTA)
    ((2, 320, 6), (1, 7, 2), 0, 1, False),     # y <320> 16*20
    ((1024, 2, 18), (7, 1, 3), 0, 1, True),    # x <1024>: forward*spectrum*inverse as two radix-32 stages (512 threads)
    ((2, 1024, 20), (1, 5, 3), 0, 1, False),   # y <1024> 16*8*8: persistent kernel on half-line tiles (8 frequencies), 11 frequencies = 1 full + 1 partial tile
    ((3, 1130, 12), (1, 23, 2), 1, 1, True),   # y <1152> 8*12*12 half-line tiles, same mode (rows beyond 1130 zero filled, crop offset 11)
    ((192, 3, 4), (9, 2, 1), 0, 2, True),      # x <192> 12*16 (guarded second stage)
    ((180, 3, 4), (9, 2, 1), 0, 1, True),      # x <180> 10*18: the 8-rank slab of the 1024-plane volume (128 + 50 halo planes)
    ((2, 180, 6), (1, 7, 2), 0, 1, False),     # y <180>, persistent pipelined form
    ((3, 5, 800), (2, 2, 9), 0, 1, True),      # rows <800> 8*10*10, pipelined, one CTA per SM
    ((2, 6, 840), (1, 3, 25), 1, 1, True),     # rows <864> 8*9*12: same mode 840 -> 864, no staging pipeline
    ((3, 5, 816), (2, 2, 49), 1, 1, True),     # rows <840> 8*7*15: same mode 816 -> 840, the longest pipelined row length
    ((3, 5, 416), (2, 2, 33), 1, 1, True),     # rows <432>: same mode 416 -> 432, forward 18*24 (384 threads), inverse 6*6*12 (576 threads)
    ((3, 5, 276), (2, 2, 25), 1, 2, False),    # rows <288>: forward 16*18, inverse 18*16
    ((320, 3, 4), (9, 2, 1), 0, 1, True),      # x <320> 16*20 as the one-tile-per-CTA walk (4-rank slab of the 1024-plane volume)
    ((288, 3, 4), (9, 2, 1), 0, 1, False),     # x <288> 16*18, same walk
]


@pytest.mark.parametrize("shape,kshape,boundary,T,den", FAST_CASES)
@pytest.mark.parametrize("variant", ["pipe", "nopipe", "generic"])
def test_fast_and_pipelined_kernels(lib, monkeypatch, shape, kshape, boundary, T, den, variant):
    if variant == "generic" and shape[0] * shape[1] * shape[2] > 40000:
        pytest.skip("generic engine already covered on small shapes")
    monkeypatch.setenv("PVD_FORCE_GENERIC", "1" if variant == "generic" else "0")
    rng = np.random.default_rng(abs(hash((shape, kshape, boundary))) % 2**32)
    maps = [rng.uniform(0, 1e3, shape) for _ in range(T)]
    maps[0][tuple(s // 2 for s in shape)] = 2e6
    w = [0.25, 0.75][:T] if T > 1 else None
    k = rng.uniform(0, 1, kshape)
    rho = rng.uniform(0.2, 2.0, shape) if den else None
    p = EmuConv(lib, shape, kshape, boundary, algo=3 if variant == "nopipe" else 0)  # 3 = PVD_ALGO_FFT_UNPIPELINED
    p.set_kernel(k)
    got = p.execute(maps, w, rho)
    p.close()
    f = lambda x: np.asarray(x, np.float32).astype(np.float64)
    acc = f(maps[0]) if T == 1 else sum(np.float64(np.float32(wi)) * f(m) for wi, m in zip(w, maps))
    ref = orc.conv_reference(acc, f(k)) if boundary == 0 else orc.conv_same(acc, f(k))
    if den:
        ref = orc.density_correct(ref, f(rho))
    assert orc.rel_err_of_peak(got, ref) <= TOL


# ---- direct tiled convolution (TMA halo tiles; the box load is emulated as a zero-filled gather) ----
@pytest.mark.parametrize("shape,kshape,den", [((9, 10, 68), (3, 3, 3), True), ((16, 8, 64), (5, 5, 5), False),
                                              ((7, 13, 132), (5, 3, 7), True), ((3, 3, 4), (7, 9, 3), True),
                                              ((8, 8, 8), (4, 2, 5), True)])
def test_direct_conv_vs_oracle(lib, shape, kshape, den):
    rng = np.random.default_rng(abs(hash((shape, kshape))) % 2**32)
    a = rng.uniform(0, 1e3, shape)
    k = rng.uniform(0, 1, kshape)
    plan = lib.plan_create(shape, kshape, 1, 2)  # same boundary, PVD_ALGO_DIRECT
    info = lib.plan_info(plan)
    assert info.algo == 2 and info.passes == 1
    nb = lib.plan_workspace_bytes(plan)
    raw = np.zeros(nb + 256, np.uint8)
    off = (-raw.ctypes.data) % 256
    ws = raw[off:off + nb]
    lib.plan_set_workspace(plan, ws.ctypes.data, nb)
    k32, a32 = np.ascontiguousarray(k, np.float32), np.ascontiguousarray(a, np.float32)
    lib.plan_set_kernel(plan, k32.ctypes.data)
    rho = np.ascontiguousarray(rng.uniform(0.2, 2.0, shape), np.float32) if den else None
    out = np.full(shape, np.nan, np.float32)
    lib.conv_execute(plan, [a32.ctypes.data], None, None if rho is None else rho.ctypes.data, 1.0, 0.1, 0.3, 2.0, out.ctypes.data)
    lib.plan_destroy(plan)
    f = lambda x: np.asarray(x, np.float32).astype(np.float64)
    ref = 2.0 * orc.conv_same(f(a), f(k))
    if den:
        ref = orc.density_correct(ref, f(rho), 1.0, 0.1, 0.3)
    assert orc.rel_err_of_peak(out, ref) <= TOL


@pytest.mark.parametrize("shape,K,den", [((16, 8, 64), 5, False), ((9, 17, 68), 3, True), ((12, 20, 72), 7, True), ((8, 8, 8), 5, True),
                                         ((7, 7, 8), 7, False)])
def test_direct_conv_circular_reference_mode_vs_oracle(lib, shape, K, den):
    """Cubic K in {3, 5, 7} in the reference's circular, origin-anchored mode (core/kernel_convolution.py:71-74):
    boundary tiles gather their halo with modulo indexing (the emulator gathers every tile)."""
    rng = np.random.default_rng(abs(hash((shape, K))) % 2**32)
    a = rng.uniform(0, 1e3, shape)
    k = rng.uniform(0, 1, (K, K, K))
    plan = lib.plan_create(shape, (K, K, K), 0, 2)  # reference boundary, PVD_ALGO_DIRECT
    info = lib.plan_info(plan)
    assert info.algo == 2 and info.passes == 1
    nb = lib.plan_workspace_bytes(plan)
    raw = np.zeros(nb + 256, np.uint8)
    off = (-raw.ctypes.data) % 256
    ws = raw[off:off + nb]
    lib.plan_set_workspace(plan, ws.ctypes.data, nb)
    k32, a32 = np.ascontiguousarray(k, np.float32), np.ascontiguousarray(a, np.float32)
    lib.plan_set_kernel(plan, k32.ctypes.data)
    rho = np.ascontiguousarray(rng.uniform(0.2, 2.0, shape), np.float32) if den else None
    out = np.full(shape, np.nan, np.float32)
    lib.conv_execute(plan, [a32.ctypes.data], None, None if rho is None else rho.ctypes.data, 1.0, 0.1, 0.0, 1.5, out.ctypes.data)
    lib.plan_destroy(plan)
    f = lambda x: np.asarray(x, np.float32).astype(np.float64)
    ref = 1.5 * orc.conv_reference(f(a), f(k))
    if den:
        ref = orc.density_correct(ref, f(rho), 1.0, 0.1, 0.0)
    assert orc.rel_err_of_peak(out, ref) <= TOL


def test_algo_selection(lib):
    from pyvoxeldosimetry_b200._capi import PvdoseError

    for shape, kshape, boundary, want in [((8, 8, 8), (5, 5, 5), 1, 2), ((8, 8, 8), (5, 5, 5), 0, 2), ((8, 8, 8), (7, 7, 7), 1, 1),
                                          ((8, 8, 6), (3, 3, 3), 1, 1), ((8, 8, 8), (5, 3, 5), 0, 1), ((4, 8, 8), (5, 5, 5), 0, 1)]:
        pl = lib.plan_create(shape, kshape, boundary, 0)
        assert lib.plan_info(pl).algo == want
        lib.plan_destroy(pl)
    with pytest.raises(PvdoseError) as e:  # the circular mode exists for cubic kernels only
        lib.plan_create((8, 8, 8), (5, 3, 5), 0, 2)
    assert e.value.code == -5


def test_batch_execute_matches_single_calls_and_the_oracle(lib):
    """pvd_conv_execute_batch (SURVEY section 8b `batch`): B volume sets, per-volume density or none, one call."""
    from pyvoxeldosimetry_b200._capi import PvdoseError

    rng = np.random.default_rng(11)
    shape, kshape, B, T = (6, 10, 12), (3, 5, 4), 3, 2
    p = EmuConv(lib, shape, kshape, 0)
    k = rng.uniform(0, 1, kshape)
    p.set_kernel(k)
    vols = [[np.ascontiguousarray(rng.uniform(0, 1e3, shape), np.float32) for _ in range(T)] for _ in range(B)]
    dens = [np.ascontiguousarray(rng.uniform(0.2, 2.0, shape), np.float32), None, np.ascontiguousarray(rng.uniform(0.2, 2.0, shape), np.float32)]
    w = [0.5, 1.5]
    outs = [np.full(shape, np.nan, np.float32) for _ in range(B)]
    lib.conv_execute_batch(p.plan, [[a.ctypes.data for a in v] for v in vols], w, [None if d is None else d.ctypes.data for d in dens],
                           1.0, 0.1, 0.0, 2.0, [o.ctypes.data for o in outs])
    f = lambda x: np.asarray(x, np.float32).astype(np.float64)
    for b in range(B):
        single = p.execute(vols[b], w, dens[b], scale=2.0)
        assert np.array_equal(outs[b], single)
        ref = 2.0 * orc.conv_reference(sum(np.float64(np.float32(wi)) * f(a) for wi, a in zip(w, vols[b])), f(np.float32(k)))
        if dens[b] is not None:
            ref = orc.density_correct(ref, f(dens[b]))
        assert orc.rel_err_of_peak(outs[b], ref) <= TOL
    lib.conv_execute_batch(p.plan, [], None, None, 1.0, 0.1, 0.0, 1.0, [])  # empty batch: nothing to do
    with pytest.raises(ValueError):
        lib.conv_execute_batch(p.plan, [[vols[0][0].ctypes.data], [vols[1][0].ctypes.data, vols[1][1].ctypes.data]], None, None, 1.0, 0.1, 0.0, 1.0,
                               [outs[0].ctypes.data, outs[1].ctypes.data])
    with pytest.raises(PvdoseError):
        lib.conv_execute_batch(p.plan, [[vols[0][0].ctypes.data]], None, None, 1.0, 0.1, 0.0, 1.0, [None])
    p.close()
