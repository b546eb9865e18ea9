"""KernelFactory - same call surface as the reference (data/dose_kernels/kernel_factory.py:12-75):
``KernelFactory().get_kernel(nuclide, tissue_type, voxel_size=1.0, grid_size=None, force_regenerate=False)``.

Repairs (SURVEY.md section 8b deltas 4, 7): the cache key includes spacing and grid (the reference keys on
nuclide+tissue only, kernel_factory.py:66, and silently returns stale kernels); the cache is in memory
(optionally a user directory), never inside the installed package; kernels are evaluated on the GPU.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from ..nuclides import NUCLIDES, REFERENCE_BROKEN
from .generators import GENERATORS, BaseKernelGenerator, _spacing3


class KernelFactory:
    _generators: Dict[str, type] = GENERATORS
    _default_grid_sizes: Dict[str, Tuple[int, int, int]] = {k: v["default_grid"] for k, v in NUCLIDES.items()}

    def __init__(self, cache_dir: Optional[str] = None):
        self.cache_dir = Path(cache_dir) if cache_dir else None
        if self.cache_dir is not None:
            self.cache_dir.mkdir(parents=True, exist_ok=True)
        self._mem: Dict[tuple, object] = {}

    @classmethod
    def supported(cls):
        return list(cls._generators.keys())

    def _check(self, nuclide: str):
        if nuclide not in self._generators:
            extra = " (registered by the reference but not computable there either)" if nuclide in REFERENCE_BROKEN else ""
            raise ValueError(f"Unsupported nuclide: {nuclide}{extra}. Supported: {self.supported()}")

    def _key(self, nuclide, tissue_type, voxel_size, grid_size):
        sp = _spacing3(voxel_size)
        grid = tuple(int(g) for g in (grid_size if grid_size is not None else self._default_grid_sizes[nuclide]))
        return (nuclide, tissue_type, sp, grid)

    def get_kernel_device(self, nuclide: str, tissue_type: str, voxel_size=1.0, grid_size: Optional[Sequence[int]] = None,
                          force_regenerate: bool = False, device=None):
        """Float32 kernel resident on the GPU (what the convolution consumes)."""
        self._check(nuclide)
        key = self._key(nuclide, tissue_type, voxel_size, grid_size) + (str(device),)
        if not force_regenerate and key in self._mem:
            return self._mem[key]
        gen: BaseKernelGenerator = self._generators[nuclide](tissue_type)
        k = gen.generate_kernel_device(key[2], key[3], device)
        self._mem[key] = k
        return k

    def get_kernel(self, nuclide: str, tissue_type: str, voxel_size=1.0, grid_size: Optional[Sequence[int]] = None,
                   force_regenerate: bool = False) -> np.ndarray:
        self._check(nuclide)
        key = self._key(nuclide, tissue_type, voxel_size, grid_size)
        path = None
        if self.cache_dir is not None:
            tag = "_".join(f"{v:g}" for v in key[2]) + "_" + "x".join(map(str, key[3]))
            path = self.cache_dir / f"{nuclide}_{tissue_type}_{tag}_kernel.npy"
            if not force_regenerate and path.exists():
                return np.load(path)
        k = self.get_kernel_device(nuclide, tissue_type, voxel_size, grid_size, force_regenerate).cpu().numpy().astype(np.float64)
        if path is not None:
            np.save(path, k)
        return k
