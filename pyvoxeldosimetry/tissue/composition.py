from pyvoxeldosimetry_b200.tissue.composition import *  # noqa: F401,F403
from pyvoxeldosimetry_b200.tissue.composition import TissueComposition  # noqa: F401
