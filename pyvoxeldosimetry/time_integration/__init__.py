from pyvoxeldosimetry_b200.time_integration import *  # noqa: F401,F403
from pyvoxeldosimetry_b200.time_integration import __all__  # noqa: F401
