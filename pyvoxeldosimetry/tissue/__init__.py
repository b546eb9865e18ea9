from pyvoxeldosimetry_b200.tissue import *  # noqa: F401,F403
from pyvoxeldosimetry_b200.tissue import __all__  # noqa: F401
