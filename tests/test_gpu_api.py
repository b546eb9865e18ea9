"""GPU tests of the drop-in Python API (reference-facing surface) and the elementwise CUDA ops."""
import os

import numpy as np
import pytest

from oracle import dose_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-4
GOLD = os.path.join(os.path.dirname(__file__), "golden")
f64 = lambda x: np.asarray(x, np.float32).astype(np.float64)


def test_device_kernel_generators_match_reference_golden():
    from pyvoxeldosimetry_b200.data.dose_kernels import KernelFactory, Lu177KernelGenerator, Y90KernelGenerator

    z = np.load(os.path.join(GOLD, "kernels_ref.npz"))
    for key in z.files:
        nuc, tissue, vox, grid = key.split("|")
        grid = tuple(int(g) for g in grid.split("x"))
        gen = (Y90KernelGenerator if nuc == "Y90" else Lu177KernelGenerator)(tissue)
        k = gen.generate_kernel(float(vox), grid)
        ref = z[key]
        ok = np.isfinite(ref)
        np.testing.assert_allclose(k[ok], ref[ok], rtol=2e-7)
        assert np.isfinite(k).all()
    f = KernelFactory()
    k1 = f.get_kernel("Y90", "water", 1.0, (15, 15, 15))
    k2 = f.get_kernel("Y90", "water", 2.0, (15, 15, 15))     # the reference would return the stale 1 mm kernel
    assert not np.allclose(k1, k2)
    ka = f.get_kernel("Lu177", "water", (1.0, 2.0, 4.0), (9, 9, 9))
    np.testing.assert_allclose(ka, orc.lu177_kernel((1.0, 2.0, 4.0), (9, 9, 9)), rtol=2e-7)


def test_calculator_reference_semantics_c1_example():
    """examples/single_timepoint_y90_physical_decay.py geometry through the calculator API."""
    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator

    calc = KernelConvolutionCalculator("Y90", "water", 1.0)
    assert calc.kernel.shape == (64, 64, 64) and calc.kernel[32, 32, 32] == 1.0
    a = orc.sphere_activity()
    d = calc.calculate_dose_rate(activity_map=a, voxel_size=(1.0, 1.0, 1.0))
    assert d.shape == a.shape and d.dtype == np.float32
    ref = orc.conv_reference(a, f64(calc.kernel))
    assert orc.rel_err_of_peak(d, ref) <= TOL
    assert np.unravel_index(d.argmax(), d.shape) == (8, 8, 8)
    # list voxel_size is accepted (the reference's tuple compare sends it into the resample stub)
    d2 = calc.calculate_dose_rate(a, [1.0, 1.0, 1.0])
    assert np.array_equal(d, d2)
    # assigning the public kernel attribute rebuilds the cached spectrum
    rng = np.random.default_rng(0)
    k = rng.uniform(0, 1, (5, 7, 3))
    calc.kernel = k
    d3 = calc.calculate_dose_rate(a, (1.0, 1.0, 1.0))
    assert orc.rel_err_of_peak(d3, orc.conv_reference(a, f64(k))) <= TOL
    with pytest.raises(Exception):
        calc.kernel = orc.y90_kernel(1.0, (9, 9, 9), centre="reference")
        calc.calculate_dose_rate(a, (1.0, 1.0, 1.0))


def test_calculator_absorbed_dose_and_density_and_float64():
    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator

    rng = np.random.default_rng(1)
    shape = (40, 44, 36)
    calc = KernelConvolutionCalculator("Lu177", "water", 4.8, config={"kernel_grid": (11, 11, 11), "output_dtype": "float64"})
    maps = [rng.uniform(0, 1e3, shape) for _ in range(4)]
    times = [4.0, 24.0, 96.0, 168.0]
    D = calc.calculate_absorbed_dose(maps, times, (4.8, 4.8, 4.8))
    assert D.dtype == np.float64
    ref = orc.absorbed_dose_trapezoid([f64(m) for m in maps], times, f64(calc.kernel))
    assert orc.rel_err_of_peak(D, ref) <= TOL
    rho = rng.choice([0.26, 1.04, 1.42], size=shape)
    Dr = calc.calculate_absorbed_dose(maps, times, (4.8, 4.8, 4.8), tissue_densities=rho)
    assert orc.rel_err_of_peak(Dr, orc.density_correct(ref, f64(rho))) <= TOL
    # anisotropic voxels: kernel evaluated on the image grid (A10) instead of the resample stub crash
    d = calc.calculate_dose_rate(maps[0], (4.8, 2.4, 1.2))
    kan = orc.lu177_kernel((4.8, 2.4, 1.2), (11, 11, 11))
    assert orc.rel_err_of_peak(d, orc.conv_reference(f64(maps[0]), f64(kan))) <= TOL
    strict = KernelConvolutionCalculator("Lu177", "water", 4.8, config={"kernel_grid": (11, 11, 11), "strict_reference": True})
    with pytest.raises(NotImplementedError):
        strict.calculate_dose_rate(maps[0], (4.8, 2.4, 1.2))


def test_many_timepoints_fold():
    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator

    rng = np.random.default_rng(2)
    calc = KernelConvolutionCalculator("Y90", "water", 1.0, config={"kernel_grid": (5, 5, 5), "boundary": "same"})
    maps = [rng.uniform(0, 1, (12, 10, 16)) for _ in range(37)]
    times = list(np.cumsum(rng.uniform(0.5, 3.0, 37)))
    D = calc.calculate_absorbed_dose(maps, times, (1.0, 1.0, 1.0))
    w = orc.trapezoid_weights(times, 3600.0)
    ref = orc.conv_same(sum(wi * f64(m) for wi, m in zip(w, maps)), f64(calc.kernel))
    assert orc.rel_err_of_peak(D, ref) <= TOL


def test_dose_calculator_front_door_modes():
    from pyvoxeldosimetry_b200 import DoseCalculator

    rng = np.random.default_rng(3)
    shape = (24, 20, 28)
    dc = DoseCalculator("Lu177", "kernel", {"kernel_resolution": 2.0, "kernel_grid": (9, 9, 9)})
    assert dc.activity_sampler.half_life == 161.52            # default from the nuclide table, not 0.0 -> ValueError
    k = f64(dc.calculator.kernel)
    maps = [rng.uniform(0, 1e3, shape) for _ in range(3)]
    t = [1.0, 24.0, 72.0]
    vs = (2.0, 2.0, 2.0)
    r = dc.calculate_dose(activity_maps=maps, time_points=t, voxel_size=vs)               # integration_mode="activity"
    acc = orc.integrate_activity_trapezoid([f64(m) for m in maps], t)
    assert r.metadata["mode"] == "multi_timepoint_activity" and len(r.dose_rate_maps) == 3
    assert orc.rel_err_of_peak(r.absorbed_dose, orc.conv_reference(acc, k)) <= TOL
    r = dc.calculate_dose(activity_maps=maps, time_points=t, voxel_size=vs, integration_mode="dose_rate")
    assert len(r.dose_rate_maps) == 3
    ref = orc.absorbed_dose_trapezoid([f64(m) for m in maps], t, k)
    assert orc.rel_err_of_peak(r.absorbed_dose, ref) <= TOL
    r = dc.calculate_dose(accumulated_activity=acc, voxel_size=vs)
    assert orc.rel_err_of_peak(r.absorbed_dose, orc.conv_reference(f64(acc), k)) <= TOL
    r = dc.calculate_dose(activity_maps=[maps[0]], time_points=[2.0], voxel_size=vs)
    rate = orc.conv_reference(f64(maps[0]), k)
    assert orc.rel_err_of_peak(r.dose_rate_maps[0], rate) <= TOL
    assert orc.rel_err_of_peak(r.absorbed_dose, rate * 161.52 * 3600 / np.log(2)) <= TOL
    rho = rng.choice([0.26, 1.04], size=shape)
    r = dc.calculate_dose(activity_maps=[maps[0]], time_points=[2.0], voxel_size=vs, tissue_densities=rho)
    assert orc.rel_err_of_peak(r.dose_rate_maps[0], orc.density_correct(rate, f64(rho))) <= TOL
    with pytest.raises(ValueError, match="length 3"):
        dc.calculate_dose(activity_maps=maps, time_points=t, voxel_size=(1.0, 1.0))
    with pytest.raises(ValueError, match="must match number of time points"):
        dc.calculate_dose(activity_maps=maps, time_points=[1.0, 2.0], voxel_size=vs)
    with pytest.raises(ValueError, match="same dimensions"):
        dc.calculate_dose(activity_maps=[maps[0], maps[1][:-1]], time_points=[1.0, 2.0], voxel_size=vs)
    with pytest.raises(ValueError, match="Invalid input"):
        dc.calculate_dose(voxel_size=vs)
    with pytest.raises(ValueError, match="Unknown integration_mode"):
        dc.calculate_dose(activity_maps=maps, time_points=t, voxel_size=vs, integration_mode="x")
    strict = DoseCalculator("Lu177", "kernel", {"kernel_resolution": 2.0, "kernel_grid": (9, 9, 9), "strict_reference": True})
    assert strict.calculate_dose(activity_maps=[maps[0]], time_points=[2.0], voxel_size=vs).absorbed_dose is None


def test_sampler_accumulation_and_monoexp_and_hu():
    import torch

    from pyvoxeldosimetry_b200 import ActivitySampler, TimeCurveFitting
    from pyvoxeldosimetry_b200.tissue import hu_to_density

    z = np.load(os.path.join(GOLD, "conv_ref.npz"))
    maps, times, A = z["tp|maps"], z["tp|times"], z["tp|A"]
    s = ActivitySampler(161.52)
    got = s.integrate_activity(list(maps), list(times))
    assert orc.rel_err_of_peak(got, A) <= 1e-6
    rates = [np.full((4, 4, 4), 1.0), np.full((4, 4, 4), 0.5)]
    got = s.integrate_dose_rates(rates, [0.0, 10.0], integration_limit=400.0)
    np.testing.assert_allclose(got, orc.integrate_dose_rates(rates, [0.0, 10.0], 400.0, 161.52), rtol=1e-6)
    tcf = TimeCurveFitting(161.52)
    params = z["a11|params"]
    np.testing.assert_allclose(tcf._calculate_accumulated_dose(params), z["a11|acc"], rtol=2e-6)
    np.testing.assert_allclose(tcf._calculate_accumulated_dose(params, 72.0), z["a11|acc72"], rtol=2e-6)
    rng = np.random.default_rng(4)
    hu = rng.integers(-1100, 3200, (20, 30, 10)).astype(np.int16)
    np.testing.assert_allclose(hu_to_density(hu), orc.hu_to_density(hu), rtol=2e-6)
    t = torch.from_numpy(hu.astype(np.float32)).cuda()
    assert hu_to_density(t).is_cuda


def test_device_tensor_in_device_tensor_out():
    import torch

    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator

    calc = KernelConvolutionCalculator("Y90", "water", 1.0, config={"kernel_grid": (7, 7, 7)})
    a = torch.rand((32, 32, 32), device="cuda")
    d = calc.calculate_dose_rate(a, (1.0, 1.0, 1.0))
    assert d.is_cuda and d.shape == a.shape
    ref = orc.conv_reference(f64(a.cpu().numpy()), f64(calc.kernel))
    assert orc.rel_err_of_peak(d.cpu().numpy(), ref) <= TOL
    pinned = torch.empty((32, 32, 32)).pin_memory()
    out = calc.calculate_dose_rate(a.cpu().numpy(), (1.0, 1.0, 1.0), out=pinned)
    assert orc.rel_err_of_peak(out, ref) <= TOL


def test_size_independent_properties_full_c3():
    """Full BASELINE size (512x512x400, 51^3): properties that need no CPU FFT of that size."""
    import torch

    from pyvoxeldosimetry_b200.engine import ConvPlan

    dev = torch.device("cuda:0")
    shape, ks = (512, 512, 400), (51, 51, 51)
    k = torch.from_numpy(orc.y90_kernel(1.0, ks).astype(np.float32)).to(dev)
    plan = ConvPlan(shape, ks, "reference", dev)
    plan.set_kernel(k)
    # delta response = rolled kernel
    a = torch.zeros(shape, device=dev)
    p = (500, 17, 390)
    a[p] = 3.0
    d = plan.execute([a])
    ref = torch.zeros(shape, device=dev)
    ref[:51, :51, :51] = 3.0 * k
    ref = torch.roll(ref, p, (0, 1, 2))
    assert float((d - ref).abs().max() / ref.abs().max()) <= TOL
    # conservation and linearity on random data
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.rand(shape, device=dev, generator=g)
    y = torch.rand(shape, device=dev, generator=g)
    dx, dy = plan.execute([x]), plan.execute([y])
    assert abs(float(dx.double().sum()) / (float(x.double().sum()) * float(k.double().sum())) - 1) < 1e-5
    dxy = plan.execute([x, y], [2.0, -3.0])
    assert float((dxy - (2 * dx - 3 * dy)).abs().max() / dx.abs().max()) <= TOL
    plan.close()


def test_pipelined_batch_api_matches_single_calls():
    import torch

    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator

    rng = np.random.default_rng(9)
    shape = (48, 40, 64)
    calc = KernelConvolutionCalculator("Y90", "water", 1.0, config={"kernel_grid": (9, 9, 9), "boundary": "same"})
    vols = [rng.uniform(0, 1e3, shape).astype(np.float32) for _ in range(5)]
    dens = [rng.choice([0.26, 1.04, 1.42], size=shape).astype(np.float32) for _ in range(5)]
    got = calc.calculate_dose_rate_batch(vols, (1.0, 1.0, 1.0), dens)
    for v, d, g in zip(vols, dens, got):
        ref = calc.calculate_dose_rate(v, (1.0, 1.0, 1.0), tissue_densities=d)
        assert np.array_equal(g, ref)
    shared = calc.calculate_dose_rate_batch(vols[:3], (1.0, 1.0, 1.0), dens[0])
    assert np.array_equal(shared[1], calc.calculate_dose_rate(vols[1], (1.0, 1.0, 1.0), tissue_densities=dens[0]))
    outs = [torch.empty(shape).pin_memory() for _ in range(3)]
    r = calc.calculate_dose_rate_batch(vols[:3], (1.0, 1.0, 1.0), None, outs)
    assert orc.rel_err_of_peak(r[2], orc.conv_same(f64(vols[2]), f64(calc.kernel))) <= TOL
    assert calc.calculate_dose_rate_batch([], (1.0, 1.0, 1.0)) == []


def test_plan_batch_execute_matches_single_calls():
    """ConvPlan.execute_batch -> pvd_conv_execute_batch: one C-ABI call for a list of volume sets (menu size: the persistent,
    dependent-launch kernels chain across volumes), bit-identical to per-volume execute()."""
    import torch

    from pyvoxeldosimetry_b200.engine import ConvPlan

    dev = torch.device("cuda:0")
    shape, ks, B, T = (64, 256, 256), (9, 9, 9), 4, 3
    g = torch.Generator(device=dev).manual_seed(3)
    plan = ConvPlan(shape, ks, "reference", dev)
    k = torch.rand(ks, device=dev, generator=g)
    plan.set_kernel(k)
    vols = [[torch.rand(shape, device=dev, generator=g) for _ in range(T)] for _ in range(B)]
    dens = [torch.rand(shape, device=dev, generator=g) + 0.5 if b % 2 == 0 else None for b in range(B)]
    w = [0.25, 1.0, 0.5]
    outs = plan.execute_batch(vols, w, dens, scale=3.0)
    for b in range(B):
        single = plan.execute(vols[b], w, dens[b], scale=3.0)
        assert torch.equal(outs[b], single)
    acc = sum(np.float64(np.float32(wi)) * a.cpu().numpy().astype(np.float64) for wi, a in zip(w, vols[1]))
    ref = 3.0 * orc.conv_reference_fast(acc, k.cpu().numpy().astype(np.float64))
    assert orc.rel_err_of_peak(outs[1].cpu().numpy(), ref) <= 1e-4
    plan.check_device_errors()
    plan.close()


def test_ct_hu_input_matches_oracle_and_density_route():
    """Density correction from the CT itself (int16 Hounsfield units, converted on the device) equals the route
    through host-side densities and the float64 oracle (conv + HU table + density correction)."""
    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator

    rng = np.random.default_rng(17)
    shape = (40, 48, 56)
    calc = KernelConvolutionCalculator("Y90", "water", 1.0, config={"kernel_grid": (9, 9, 9), "boundary": "same"})
    vols = [rng.uniform(0, 1e3, shape).astype(np.float32) for _ in range(4)]
    cts = [rng.integers(-1000, 1500, size=shape).astype(np.int16) for _ in range(4)]
    k64 = f64(calc.kernel)
    got = calc.calculate_dose_rate(vols[0], (1.0, 1.0, 1.0), ct_hu=cts[0])
    ref = orc.density_correct(orc.conv_same(f64(vols[0]), k64), orc.hu_to_density(cts[0]))
    assert orc.rel_err_of_peak(got, ref) <= TOL
    batch = calc.calculate_dose_rate_batch(vols, (1.0, 1.0, 1.0), ct_hu=cts)
    for v, c, b in zip(vols, cts, batch):
        assert np.array_equal(b, calc.calculate_dose_rate(v, (1.0, 1.0, 1.0), ct_hu=c))
    shared = calc.calculate_dose_rate_batch(vols[:2], (1.0, 1.0, 1.0), ct_hu=cts[1])
    assert np.array_equal(shared[0], calc.calculate_dose_rate(vols[0], (1.0, 1.0, 1.0), ct_hu=cts[1]))
    dose = calc.calculate_absorbed_dose(vols[:3], [1.0, 5.0, 20.0], (1.0, 1.0, 1.0), ct_hu=cts[2].astype(np.float32))
    w = [2.0 * 3600, (2.0 + 7.5) * 3600, 7.5 * 3600]
    acc = sum(wi * f64(v) for wi, v in zip(w, vols[:3]))
    assert orc.rel_err_of_peak(dose, orc.density_correct(orc.conv_same(acc, k64), orc.hu_to_density(cts[2]))) <= TOL
    with pytest.raises(ValueError):
        calc.calculate_dose_rate(vols[0], (1.0, 1.0, 1.0), tissue_densities=np.ones(shape, np.float32), ct_hu=cts[0])
    with pytest.raises(ValueError):
        calc.calculate_dose_rate(vols[0], (1.0, 1.0, 1.0), ct_hu=cts[0][:-1])
