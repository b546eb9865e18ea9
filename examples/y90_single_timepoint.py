"""Single-timepoint Y-90 dose on a voxelised sphere through the drop-in API (`pyvoxeldosimetry` alias).
Same scenario as the reference's examples/single_timepoint_y90_physical_decay.py; prints numbers instead
of plotting.  Needs a CUDA device."""
import numpy as np

from pyvoxeldosimetry.core import DoseCalculator

n, c, radius = 48, 24, 8
g = np.arange(n) - c
activity = np.where(g[:, None, None] ** 2 + g[None, :, None] ** 2 + g[None, None, :] ** 2 <= radius ** 2, 2e6, 0.0)
calc = DoseCalculator("Y90", method="kernel", config={"kernel_resolution": 1.0, "half_life": 64.1, "time_units": "hours"})
res = calc.calculate_dose(activity_maps=[activity], time_points=[2.0], voxel_size=(1.0, 1.0, 1.0))
rate, dose = res.dose_rate_maps[0], res.absorbed_dose
print(f"total activity          : {activity.sum():.3e} Bq")
print(f"peak dose rate          : {rate.max():.6e} at voxel {np.unravel_index(rate.argmax(), rate.shape)}")
print(f"peak absorbed dose      : {dose.max():.6e} (physical decay from the scan time)")
same = DoseCalculator("Y90", "kernel", {"half_life": 64.1, "boundary": "same", "kernel_grid": (51, 51, 51)})
r2 = same.calculate_dose(activity_maps=[activity], time_points=[2.0], voxel_size=(1.0, 1.0, 1.0)).dose_rate_maps[0]
print(f"'same' boundary peak    : {r2.max():.6e} at voxel {np.unravel_index(r2.argmax(), r2.shape)} (centred, zero boundary)")
