from pyvoxeldosimetry_b200.core.utils import calculate_dvh, interpolate_timepoints, interpolation_weights, load_dose_map, load_kernel, save_dose_map, save_kernel  # noqa: F401
