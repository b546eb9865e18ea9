"""Drop-in alias: `import pyvoxeldosimetry` resolves to the B200-native kernel-convolution path
(pyvoxeldosimetry_b200).  Only the names on that path exist; Monte-Carlo, GATE, segmentation, IO and
registration of the reference are outside this package."""
from pyvoxeldosimetry_b200 import *  # noqa: F401,F403
from pyvoxeldosimetry_b200 import __all__, __version__  # noqa: F401
