from pyvoxeldosimetry_b200.io import *  # noqa: F401,F403
from pyvoxeldosimetry_b200.io import __all__  # noqa: F401
