from .composition import METAL_THRESHOLD_HU, TissueComposition
from .density import HU_KNOTS, hu_to_density

__all__ = ["HU_KNOTS", "hu_to_density", "TissueComposition", "METAL_THRESHOLD_HU"]
