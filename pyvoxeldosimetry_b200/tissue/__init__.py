from .density import HU_KNOTS, hu_to_density

__all__ = ["HU_KNOTS", "hu_to_density"]
