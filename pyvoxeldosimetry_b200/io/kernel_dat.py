"""Binary `.dat` dose-voxel-kernel files (reference core/utils.py:17-51 `load_kernel`).

Layout read by the reference (little-endian, 80-byte header):
    int32[3] dims | float32 voxel_size | float32 total_energy | float32 scaling | int32[6] timestamp
    (Y, M, D, h, m, s) | int8[32] user | float32 data[prod(dims)]  (C order), kernel = data * scaling.
The reference ships no writer; `save_kernel` writes exactly what `load_kernel` reads.  (The reference's own
documentation and its verify_file_integrity disagree about the header size - 24 / 44 bytes, utils.py:283-286,
data/dose_kernels/README.md:326-341 - `load_kernel` is the format that is actually parsed, so it wins.)
"""
from __future__ import annotations

import struct
from datetime import datetime
from pathlib import Path
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np

HEADER_BYTES = 80


def load_kernel(filename: Union[str, Path]) -> Tuple[np.ndarray, Dict[str, Any]]:
    raw = Path(filename).read_bytes()
    if len(raw) < HEADER_BYTES:
        raise ValueError(f"{filename}: too short for a kernel header ({len(raw)} < {HEADER_BYTES} bytes)")
    dims = np.frombuffer(raw, dtype="<i4", count=3, offset=0)
    voxel_size, total_energy, scaling = struct.unpack_from("<3f", raw, 12)
    timestamp = np.frombuffer(raw, dtype="<i4", count=6, offset=24)
    user = raw[48:80]
    if np.any(dims <= 0):
        raise ValueError(f"{filename}: invalid kernel dimensions {dims.tolist()}")
    count = int(np.prod(dims.astype(np.int64)))
    data = np.frombuffer(raw, dtype="<f4", offset=HEADER_BYTES)
    if data.size != count:
        raise ValueError(f"{filename}: {data.size} float32 values after the header, dimensions {dims.tolist()} need {count}")
    kernel = data.reshape(tuple(int(d) for d in dims)) * np.float32(scaling)
    metadata = {
        "dimensions": dims.copy(),
        "voxel_size": np.float32(voxel_size),
        "total_energy": np.float32(total_energy),
        "scaling_factor": np.float32(scaling),
        "creation_date": datetime(*[int(v) for v in timestamp]).strftime("%Y-%m-%d %H:%M:%S"),
        "created_by": bytes(user).decode().strip("\x00"),
    }
    return kernel, metadata


def save_kernel(filename: Union[str, Path], kernel: np.ndarray, voxel_size: float, total_energy: float = 0.0,
                scaling: float = 1.0, created: Optional[datetime] = None, user: str = "") -> None:
    """Write `kernel` so that load_kernel returns it: the stored samples are kernel / scaling (float32)."""
    k = np.asarray(kernel)
    if k.ndim != 3:
        raise ValueError("kernel must be 3-D")
    if scaling == 0:
        raise ValueError("scaling must be non-zero")
    created = created or datetime.now()
    u = user.encode()[:32].ljust(32, b"\x00")
    head = struct.pack("<3i3f6i", *k.shape, float(voxel_size), float(total_energy), float(scaling), created.year,
                       created.month, created.day, created.hour, created.minute, created.second) + u
    assert len(head) == HEADER_BYTES
    data = (k.astype(np.float64) / float(scaling)).astype("<f4")
    with open(filename, "wb") as f:
        f.write(head)
        f.write(np.ascontiguousarray(data).tobytes())
