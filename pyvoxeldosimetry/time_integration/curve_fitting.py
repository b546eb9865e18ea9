from pyvoxeldosimetry_b200.time_integration.accumulation import TimeCurveFitting  # noqa: F401
