from pyvoxeldosimetry_b200.data.dose_kernels import *  # noqa: F401,F403
from pyvoxeldosimetry_b200.data.dose_kernels import __all__  # noqa: F401
