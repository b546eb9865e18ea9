"""Device timings of the section-8f kernels (CUDA events, inputs resident in HBM, larger than L2):
mono-exponential fit + integral (256^3, T = 4), CT prepare (512x512x400, 0.05 % metal voxels), DVH (512x512x400).
Prints one JSON object; achieved GB/s = algorithmic bytes / time against the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyvoxeldosimetry_b200 import engine
from pyvoxeldosimetry_b200.tissue.density import HU_KNOTS

peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(1)
res = {}

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# fit: 256^3, 4 time points (config C2 shape)
shape, times = (256, 256, 256), [4.0, 24.0, 96.0, 168.0]
lam0 = float(np.log(2) / 161.52)
A0 = torch.rand(shape, device=dev, generator=g) * 1e6 + 1e2
lam = lam0 * (0.7 + 3.3 * torch.rand(shape, device=dev, generator=g))
vols = [(A0 * torch.exp(-lam * t) * (1 + 0.05 * torch.randn(shape, device=dev, generator=g))).contiguous() for t in times]
ms = timeit(lambda: engine.monoexp_fit(vols, times, None, lam0, 100 * 161.52, want_params=False))
nb = 4 * (len(times) + 1) * A0.numel()
res['monoexp_fit_256^3_T4_integral_only'] = {'ms': ms, 'voxels_per_s': A0.numel() / ms * 1e3, 'algorithmic_bytes': nb, 'gbs': nb / ms / 1e6, 'frac_of_measured_peak': nb / ms / 1e6 / peak}
ms = timeit(lambda: engine.monoexp_fit(vols, times, None, lam0, 100 * 161.52, want_params=True))
nb = 4 * (len(times) + 3) * A0.numel()
res['monoexp_fit_256^3_T4_with_params'] = {'ms': ms, 'voxels_per_s': A0.numel() / ms * 1e3, 'algorithmic_bytes': nb, 'gbs': nb / ms / 1e6, 'frac_of_measured_peak': nb / ms / 1e6 / peak}
del vols, A0, lam
# CT prepare: 512x512x400
shape = (512, 512, 400)
hu = torch.rand(shape, device=dev, generator=g) * 2800 - 1000
hu.view(-1)[torch.randint(0, hu.numel(), (hu.numel() // 2000,), device=dev, generator=g)] = 3000.0
ranges = [(-1000, -900), (-900, -500), (-100, 100), (300, 3000), (-10, 10)]
ms = timeit(lambda: engine.ct_prepare(hu, 2000.0, HU_KNOTS, ranges, want_corrected=False), reps=10)
nb = (4 + 4 + 1) * hu.numel()
res['ct_prepare_512x512x400_rho+labels'] = {'ms': ms, 'voxels_per_s': hu.numel() / ms * 1e3, 'algorithmic_bytes': nb, 'gbs': nb / ms / 1e6, 'frac_of_measured_peak': nb / ms / 1e6 / peak, 'metal_fraction': 1 / 2000}
# DVH: 512x512x400 dose, uint8 mask
dose = torch.rand(shape, device=dev, generator=g) * 50
mask = (torch.rand(shape, device=dev, generator=g) > 0.5).to(torch.uint8)
edges = torch.linspace(0, 50, 1001, device=dev)
ms1 = timeit(lambda: engine.roi_minmax(dose, mask), reps=10)
ms2 = timeit(lambda: engine.dvh_histogram(dose, mask, edges), reps=10)
nb = 5 * dose.numel()
res['dvh_512x512x400_minmax'] = {'ms': ms1, 'algorithmic_bytes': nb, 'gbs': nb / ms1 / 1e6, 'frac_of_measured_peak': nb / ms1 / 1e6 / peak, 'note': 'includes the 16-byte D2H + stream sync'}
res['dvh_512x512x400_histogram_1000bins'] = {'ms': ms2, 'algorithmic_bytes': nb, 'gbs': nb / ms2 / 1e6, 'frac_of_measured_peak': nb / ms2 / 1e6 / peak}
del dose, mask, hu
torch.cuda.empty_cache()
# time-axis combinations (pvd_weighted_combine): the trapezoid sum of 4 volumes at the C3 size, and interpolate_timepoints
# (4 sampled 256^3 volumes -> 8 interpolated ones; linear = 2 weights per output, cubic = dense weights)
from pyvoxeldosimetry_b200.core.utils import interpolation_weights
vols = [torch.rand(shape, device=dev, generator=g) for _ in range(4)]
ms = timeit(lambda: engine.weighted_sum(vols, [0.5, 1.0, 1.0, 0.5]), reps=10)
nb = 4 * 5 * vols[0].numel()
res['weighted_sum_512x512x400_T4'] = {'ms': ms, 'algorithmic_bytes': nb, 'gbs': nb / ms / 1e6, 'frac_of_measured_peak': nb / ms / 1e6 / peak}
del vols
vols = [torch.rand((256, 256, 256), device=dev, generator=g) for _ in range(4)]
times, new = [4.0, 24.0, 96.0, 168.0], list(np.linspace(4.0, 168.0, 8))
for kind in ('linear', 'cubic'):
    W = interpolation_weights(times, new, kind).tolist()
    ms = timeit(lambda: engine.weighted_combine(vols, W), reps=20)
    nb = 4 * (4 + 8) * vols[0].numel()
    res[f'interpolate_timepoints_256^3_T4_to_8_{kind}'] = {'ms': ms, 'algorithmic_bytes': nb, 'gbs': nb / ms / 1e6, 'frac_of_measured_peak': nb / ms / 1e6 / peak}
res['peak_gbs_measured'] = peak
print(json.dumps(res))
