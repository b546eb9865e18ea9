"""GPU parity: CUDA path (through the C ABI) vs the float64 oracle.  Gate: max|d - ref| / max|ref| <= 1e-4
(BASELINE.json north_star); in practice the fp32 engine sits at ~1e-6."""
import json
import os

import numpy as np
import pytest

from oracle import dose_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-4  # north-star tolerance, relative to peak dose
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    return torch


def run_plan(torch, a_list, k, boundary, weights=None, density=None, **kw):
    from pyvoxeldosimetry_b200.engine import ConvPlan

    dev = torch.device("cuda:0")
    plan = ConvPlan(a_list[0].shape, k.shape, boundary, dev)
    plan.set_kernel(k)
    acts = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev) for a in a_list]
    den = None if density is None else torch.from_numpy(np.ascontiguousarray(density, dtype=np.float32)).to(dev)
    out = plan.execute(acts, weights, den, **kw)
    plan.check_device_errors()  # synchronises; raises if a TMA watchdog fired
    res = out.cpu().numpy()
    plan.close()
    return res


CASES = [
    ((16, 12, 20), (5, 7, 3)),
    ((10, 9, 8), (12, 4, 11)),      # kernel larger than the grid: crop
    ((7, 11, 13), (3, 3, 3)),       # prime lengths -> generic radix stages
    ((8, 8, 8), (1, 1, 1)),
    ((12, 10, 14), (4, 6, 2)),      # even kernel
    ((48, 48, 48), (64, 64, 64)),   # config C1 geometry
    ((64, 80, 50), (9, 9, 9)),
    ((33, 65, 127), (5, 5, 5)),     # odd / prime-ish lengths (3*11, 5*13, 127)
    ((1, 40, 36), (1, 5, 5)),       # degenerate axis
    ((100, 96, 90), (31, 31, 31)),
]


@pytest.mark.parametrize("shape,kshape", CASES)
@pytest.mark.parametrize("boundary", ["reference", "same"])
def test_conv_matches_oracle(torch_cuda, shape, kshape, boundary):
    rng = np.random.default_rng(hash((shape, kshape)) % (2**32))
    a = rng.uniform(0.0, 1e3, size=shape)
    a[tuple(s // 2 for s in shape)] = 2e6
    k = rng.uniform(0.0, 1.0, size=kshape)
    got = run_plan(torch_cuda, [a], k, boundary)
    a32, k32 = a.astype(np.float32).astype(np.float64), k.astype(np.float32).astype(np.float64)
    ref = orc.conv_reference_fast(a32, k32) if boundary == "reference" else orc.conv_same(a32, k32, fast=True)
    assert got.shape == ref.shape
    assert orc.rel_err_of_peak(got, ref) <= TOL


def test_golden_reference_vectors(torch_cuda):
    """Outputs of the REAL reference calculator (tests/golden/conv_ref.npz, made by oracle/gen_golden.py)."""
    z = np.load(os.path.join(GOLD, "conv_ref.npz"))
    names = sorted({k.split("|")[0] for k in z.files if k.endswith("|d")})
    assert names
    for name in names:
        a, k, d = z[name + "|a"], z[name + "|k"], z[name + "|d"]
        got = run_plan(torch_cuda, [a], k, "reference")
        assert orc.rel_err_of_peak(got, d) <= TOL, name


def test_golden_time_integrated(torch_cuda):
    z = np.load(os.path.join(GOLD, "conv_ref.npz"))
    maps, times, k, D = z["tp|maps"], z["tp|times"], z["tp|k"], z["tp|D"]
    w = orc.trapezoid_weights(times, 3600.0)
    got = run_plan(torch_cuda, list(maps), k, "reference", weights=list(w))
    assert orc.rel_err_of_peak(got, D) <= TOL


def test_c1_known_answers(torch_cuda):
    """Config C1 (examples/single_timepoint_y90_physical_decay.py) with the finite-centre Y90 kernel."""
    kats = json.load(open(os.path.join(GOLD, "kats.json")))
    planes = np.load(os.path.join(GOLD, "c1_ref.npz"))
    a = orc.sphere_activity()
    k = orc.y90_kernel(1.0, (64, 64, 64), "water")
    got = run_plan(torch_cuda, [a], k, "reference").astype(np.float64)
    peak = kats["c1_max"]
    assert abs(got.max() - peak) / peak <= TOL
    assert int(got.argmax()) == int(kats["c1_argmax"])
    assert abs(got[0, 0, 0] - kats["c1_d000"]) / peak <= TOL
    assert abs(got[24, 24, 24] - kats["c1_d242424"]) / peak <= TOL
    assert abs(got.sum() - kats["c1_sum"]) / kats["c1_sum"] <= TOL
    for name, sl in (("plane_x8", got[8]), ("plane_y8", got[:, 8]), ("plane_z8", got[:, :, 8])):
        assert np.max(np.abs(sl - planes[name])) / peak <= TOL


def test_density_and_scale_fused(torch_cuda):
    rng = np.random.default_rng(7)
    shape, kshape = (40, 36, 50), (7, 7, 7)
    a = rng.uniform(0, 1e3, shape)
    k = rng.uniform(0, 1, kshape)
    rho = rng.choice([0.00129, 0.26, 1.04, 1.42], size=shape)
    got = run_plan(torch_cuda, [a], k, "same", density=rho, rho_ref=1.0, rho_min=0.1, rho_cut=0.01, scale=2.5)
    a32, k32, r32 = (x.astype(np.float32).astype(np.float64) for x in (a, k, rho))
    ref = orc.density_correct(2.5 * orc.conv_same(a32, k32, fast=True), r32, 1.0, 0.1, 0.01)
    assert orc.rel_err_of_peak(got, ref) <= TOL


def test_nonfinite_kernel_rejected(torch_cuda):
    from pyvoxeldosimetry_b200._capi import PvdoseError
    from pyvoxeldosimetry_b200.engine import ConvPlan

    k = orc.y90_kernel(1.0, (9, 9, 9), "water", centre="reference")  # NaN at the centre like the reference
    plan = ConvPlan((16, 16, 16), (9, 9, 9), "reference", "cuda:0")
    with pytest.raises(PvdoseError) as e:
        plan.set_kernel(k)
    assert e.value.code == -4
    plan.close()


@pytest.mark.parametrize("boundary", ["reference", "same"])
def test_c2_lu177_256_time_integrated(torch_cuda, boundary):
    """Config C2: Lu-177, 4 time points, 256^3, 31^3 kernel @ 4.8 mm."""
    rng = np.random.default_rng(177)
    n = (256, 256, 256)
    x, y, z = np.ogrid[:256, :256, :256]
    a0 = rng.uniform(0.0, 1e3, n)
    for (c, r, v) in (((128, 128, 128), 20, 1e6), ((60, 90, 170), 8, 5e6), ((200, 50, 80), 12, 2e6)):
        a0[(x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2 <= r * r] = v
    times = [4.0, 24.0, 96.0, 168.0]
    maps = [(a0 * np.exp(-np.log(2) * t / 161.52)).astype(np.float32) for t in times]
    k = orc.lu177_kernel(4.8, (31, 31, 31), "water")
    w = orc.trapezoid_weights(times, 3600.0)
    got = run_plan(torch_cuda, maps, k, boundary, weights=list(w))
    acc = sum(wi * m.astype(np.float64) for wi, m in zip(w, maps))
    k32 = k.astype(np.float32).astype(np.float64)
    ref = orc.conv_reference_fast(acc, k32) if boundary == "reference" else orc.conv_same(acc, k32, fast=True)
    assert orc.rel_err_of_peak(got, ref) <= TOL


@pytest.mark.parametrize("shape,kshape", [((64, 64, 64), (5, 5, 5)), ((40, 36, 52), (3, 3, 3)), ((33, 65, 128), (5, 3, 7)),
                                          ((100, 96, 92), (5, 5, 5)), ((9, 7, 4), (7, 9, 3))])
def test_direct_tma_conv_matches_oracle(torch_cuda, shape, kshape):
    """PVD_ALGO_DIRECT: TMA-staged halo tiles, 'same' boundary; also AUTO must pick it for 5^3 and agree with FFT."""
    from pyvoxeldosimetry_b200.engine import ConvPlan

    torch = torch_cuda
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(abs(hash((shape, kshape))) % 2**32)
    a = rng.uniform(0.0, 1e3, size=shape).astype(np.float32)
    k = rng.uniform(0.0, 1.0, size=kshape).astype(np.float32)
    rho = rng.choice([0.26, 1.04, 1.42], size=shape).astype(np.float32)
    ref = orc.density_correct(orc.conv_same(a.astype(np.float64), k.astype(np.float64), fast=True), rho)
    outs = {}
    for name, algo in (("direct", 2), ("fft", 1)):
        plan = ConvPlan(shape, kshape, "same", dev, algo)
        assert plan.info.algo == algo
        plan.set_kernel(k)
        out = plan.execute([torch.from_numpy(a).to(dev)], None, torch.from_numpy(rho).to(dev))
        torch.cuda.synchronize()
        outs[name] = out.cpu().numpy()
        plan.close()
        assert orc.rel_err_of_peak(outs[name], ref) <= TOL, name
    # time-weighted input through the direct path (folded by weighted_sum first)
    plan = ConvPlan(shape, kshape, "same", dev, 2)
    plan.set_kernel(k)
    out = plan.execute([torch.from_numpy(a).to(dev), torch.from_numpy(a).to(dev)], [0.25, 0.5]).cpu().numpy()
    plan.close()
    assert orc.rel_err_of_peak(out, 0.75 * orc.conv_same(a.astype(np.float64), k.astype(np.float64), fast=True)) <= TOL


@pytest.mark.parametrize("shape,K", [((64, 64, 64), 5), ((40, 36, 52), 3), ((100, 96, 92), 7), ((136, 80, 200), 5), ((8, 8, 8), 7), ((512, 32, 400), 5)])
def test_direct_conv_circular_reference_mode(torch_cuda, shape, K):
    """Cubic direct kernels in the reference's circular, origin-anchored mode: interior tiles by TMA, the tiles whose halo
    crosses index 0 by a modulo gather; AUTO picks the direct path for 5^3 under the DEFAULT boundary now."""
    from pyvoxeldosimetry_b200.engine import ConvPlan

    torch = torch_cuda
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(abs(hash((shape, K))) % 2**32)
    a = rng.uniform(0.0, 1e3, size=shape).astype(np.float32)
    a[tuple(s // 3 for s in shape)] = 2e6
    k = rng.uniform(0.0, 1.0, size=(K, K, K)).astype(np.float32)
    rho = rng.choice([0.26, 1.04, 1.42], size=shape).astype(np.float32)
    ref = orc.density_correct(orc.conv_reference_fast(a.astype(np.float64), k.astype(np.float64)), rho)
    for algo in (2, 1, 0):
        plan = ConvPlan(shape, (K, K, K), "reference", dev, algo)
        want = 2 if (algo == 2 or (algo == 0 and K <= 5)) else 1
        assert plan.info.algo == want, (algo, plan.info.algo)
        plan.set_kernel(k)
        out = plan.execute([torch.from_numpy(a).to(dev)], None, torch.from_numpy(rho).to(dev), scale=1.0)
        plan.check_device_errors()
        assert orc.rel_err_of_peak(out.cpu().numpy(), ref) <= TOL, algo
        plan.close()


def test_c3_full_size_vs_oracle(torch_cuda):
    """Config C3 at its full size: 512x512x400 Y90 volume, 51^3 kernel, density-corrected, reference boundary mode,
    against the float64 oracle (scipy real transforms with all host threads: same mathematics as the literal
    np.fft expression, SURVEY section 8d "Baseline B")."""
    rng = np.random.default_rng(90)
    shape = (512, 512, 400)
    a = rng.uniform(0.0, 1e2, shape).astype(np.float32)
    a[200:300, 220:330, 150:260] = 2e6
    x = (np.arange(512, dtype=np.float32) - 256) / 215.0
    body = (x[:, None] ** 2 + (x[None, :] * 1.4) ** 2) <= 1.0
    plane = np.where(body, 1.04, 0.00129).astype(np.float32)
    plane[(np.abs(x[:, None] - 0.45) < 0.25) & (np.abs(x[None, :]) < 0.3)] = 0.26
    rho = np.ascontiguousarray(np.broadcast_to(plane[:, :, None], shape))
    k = orc.y90_kernel(1.0, (51, 51, 51), "water")
    got = run_plan(torch_cuda, [a], k, "reference", density=rho)
    ref = orc.conv_reference_fast(a.astype(np.float64), k.astype(np.float32).astype(np.float64))
    ref = orc.density_correct(ref, rho.astype(np.float64))
    assert orc.rel_err_of_peak(got, ref) <= TOL


@pytest.mark.parametrize("shape,boundary,fft_shape", [
    ((40, 1030, 830), "same", (None, 1152, 864)),      # slab sizes of the 1024x1024x800 volume: columns <1152>, rows <864>
    ((24, 1024, 800), "reference", (24, 1024, 800)),   # reference mode: columns <1024>, rows <800>
    ((270, 64, 100), "same", (320, None, None)),       # 256-plane slab + halo -> <320>
    ((150, 48, 60), "same", (180, None, None)),        # 128-plane slab + halo -> <180>
    ((165, 48, 60), "same", (192, None, None)),        # -> <192>
    ((24, 40, 400), "same", (None, None, 432)),        # 400 + kernel reach -> rows <432>: forward 18*24, inverse 6*6*12 (576 threads)
    ((24, 262, 262), "same", (None, 288, 288)),        # 256 + reach -> <288>: rows forward 16*18 / inverse 18*16, y passes <288>
    ((270, 40, 262), "same", (320, None, 288)),        # x walk <320> over rows <288>
])
def test_slab_menu_sizes_vs_oracle(torch_cuda, shape, boundary, fft_shape):
    from pyvoxeldosimetry_b200.engine import ConvPlan

    rng = np.random.default_rng(5)
    a = rng.uniform(0.0, 1e3, shape).astype(np.float32)
    a[tuple(s // 2 for s in shape)] = 2e6
    k = orc.y90_kernel(1.0, (51, 51, 51), "water") if shape[0] >= 150 else orc.y90_kernel(1.0, (25, 51, 51), "water")
    plan = ConvPlan(shape, k.shape, boundary, "cuda:0")
    for got_m, want in zip(plan.fft_shape, fft_shape):
        assert want is None or got_m == want, (plan.fft_shape, fft_shape)
    plan.close()
    got = run_plan(torch_cuda, [a], k, boundary)
    k32 = k.astype(np.float32).astype(np.float64)
    ref = orc.conv_reference_fast(a.astype(np.float64), k32) if boundary == "reference" else orc.conv_same(a.astype(np.float64), k32, fast=True)
    assert orc.rel_err_of_peak(got, ref) <= TOL
