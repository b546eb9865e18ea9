"""Lu-177 four-timepoint absorbed dose with voxel-wise density correction (config C2 of BASELINE.json at
reduced size).  One fused convolution: sum_t w_t a_t -> FFT conv -> density scale.  Needs a CUDA device."""
import numpy as np

from pyvoxeldosimetry.core import DoseCalculator

rng = np.random.default_rng(177)
n = 128
g = np.arange(n) - n // 2
a0 = rng.uniform(0, 1e3, (n, n, n))
a0[g[:, None, None] ** 2 + g[None, :, None] ** 2 + g[None, None, :] ** 2 <= 15 ** 2] = 1e6
times = [4.0, 24.0, 96.0, 168.0]
maps = [a0 * np.exp(-np.log(2) * t / 161.52) for t in times]
rho = np.full((n, n, n), 1.04, np.float32)
rho[: n // 3] = 0.26
calc = DoseCalculator("Lu177", "kernel", {"kernel_resolution": 4.8, "kernel_grid": (31, 31, 31), "boundary": "same",
                                           "return_dose_rate_maps": False})
res = calc.calculate_dose(activity_maps=maps, time_points=times, voxel_size=(4.8, 4.8, 4.8), tissue_densities=rho,
                          integration_mode="dose_rate")
print("absorbed dose max/mean:", float(res.absorbed_dose.max()), float(res.absorbed_dose.mean()))
