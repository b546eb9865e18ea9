from pyvoxeldosimetry_b200.core import *  # noqa: F401,F403
from pyvoxeldosimetry_b200.core import __all__  # noqa: F401
