from .accumulation import TimeCurveFitting

__all__ = ["TimeCurveFitting"]
