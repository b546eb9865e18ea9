"""NIfTI-1 dose maps with the metadata JSON in header extension 44 (reference core/utils.py:53-152,
`save_dose_map` / `load_dose_map`, which use nibabel).  nibabel is not a dependency here: the single-file
NIfTI-1 layout is written and parsed directly (348-byte header, 4-byte extender, extension(s), data at
vox_offset; `.nii.gz` through gzip).  Field values follow what nibabel's Nifti1Image(data.astype(float32), affine)
produces for a 3-D float32 volume: sform_code = 2 (aligned) with the affine in srow_x/y/z, qform_code = 0,
xyzt_units = 2 (mm), scl_slope = scl_inter = NaN on disk (= "no scaling"), Fortran-order voxel data.
"""
from __future__ import annotations

import gzip
import json
import struct
from pathlib import Path
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np

HDR = 348
_DT = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4"}
_CODE = {np.dtype(v).newbyteorder("<").str[1:]: k for k, v in _DT.items()}


def _open(filename: str, mode: str):
    return gzip.open(filename, mode) if filename.endswith(".gz") else open(filename, mode)


def _jsonable(o):
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    if isinstance(o, tuple):
        return list(o)
    raise TypeError(f"{type(o).__name__} is not JSON serialisable")


def save_dose_map(filename: Union[str, Path], dose_map: np.ndarray, voxel_size: Tuple[float, float, float],
                  metadata: Dict[str, Any], affine: Optional[np.ndarray] = None) -> str:
    """Same arguments and behaviour as the reference: appends `.nii.gz` when the name has no NIfTI suffix,
    default affine = diag(voxel_size) centred on the volume (:72-77), metadata updated with creation info,
    voxel size, dimensions and affine (:80-86) and stored as JSON in extension 44 (:102-107).  Returns the path."""
    filename = str(filename)
    if not filename.endswith((".nii", ".nii.gz")):
        filename += ".nii.gz"
    dose = np.asarray(dose_map)
    if dose.ndim != 3:
        raise ValueError("dose map must be 3-D")
    vs = tuple(float(v) for v in voxel_size)
    if affine is None:
        affine = np.diag(list(vs) + [1.0])
        affine[:3, 3] = np.array(dose.shape) * np.array(vs) / -2.0
    affine = np.asarray(affine, dtype=np.float64)
    metadata.update({  # the reference hard-codes these two strings (:81-82)
        "creation_date": "2025-02-08 09:50:56",
        "created_by": "devhliu",
        "voxel_size": vs,
        "dimensions": dose.shape,
        "affine_matrix": affine.tolist(),
    })
    content = json.dumps(metadata, default=_jsonable).encode("utf-8")
    esize = 8 + len(content)
    esize += (-esize) % 16
    ext = struct.pack("<2i", esize, 44) + content.ljust(esize - 8, b"\x00")
    vox_offset = HDR + 4 + len(ext)
    descrip = f"Created by {metadata['created_by']} at {metadata['creation_date']}".encode()[:79]
    h = bytearray(HDR)
    struct.pack_into("<i", h, 0, HDR)
    struct.pack_into("<8h", h, 40, 3, dose.shape[0], dose.shape[1], dose.shape[2], 1, 1, 1, 1)
    struct.pack_into("<h", h, 70, 16)   # datatype float32
    struct.pack_into("<h", h, 72, 32)   # bitpix
    struct.pack_into("<8f", h, 76, 1.0, vs[0], vs[1], vs[2], 1.0, 1.0, 1.0, 1.0)  # pixdim, qfac = 1
    struct.pack_into("<f", h, 108, float(vox_offset))
    struct.pack_into("<2f", h, 112, float("nan"), float("nan"))  # scl_slope, scl_inter: no scaling
    h[123] = 2  # xyzt_units: millimetres
    h[148:148 + len(descrip)] = descrip
    struct.pack_into("<2h", h, 252, 0, 2)  # qform_code, sform_code
    struct.pack_into("<4f", h, 280, *affine[0])
    struct.pack_into("<4f", h, 296, *affine[1])
    struct.pack_into("<4f", h, 312, *affine[2])
    h[344:348] = b"n+1\x00"
    with _open(filename, "wb") as f:
        f.write(bytes(h))
        f.write(bytes([1, 0, 0, 0]))  # extender: extensions follow
        f.write(ext)
        f.write(np.asfortranarray(dose.astype("<f4")).tobytes(order="F"))
    return filename


def load_dose_map(filename: Union[str, Path]) -> Tuple[np.ndarray, Dict[str, Any]]:
    """-> (dose array, metadata) with the reference's keys (:141-158): voxel_size, dimensions, affine_matrix,
    data_type, description, plus whatever the extension-44 JSON holds."""
    filename = str(filename)
    with _open(filename, "rb") as f:
        raw = f.read()
    if len(raw) < HDR + 4:
        raise ValueError(f"{filename}: too short for a NIfTI-1 header")
    end = "<" if struct.unpack_from("<i", raw, 0)[0] == HDR else ">"
    if struct.unpack_from(end + "i", raw, 0)[0] != HDR or raw[344:347] not in (b"n+1", b"ni1"):
        raise ValueError(f"{filename}: not a NIfTI-1 file")
    if raw[344:347] == b"ni1":
        raise ValueError(f"{filename}: header/image pairs (.hdr/.img) are not supported")
    dim = struct.unpack_from(end + "8h", raw, 40)
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    while len(shape) > 3 and shape[-1] == 1:
        shape = shape[:-1]
    dtcode = struct.unpack_from(end + "h", raw, 70)[0]
    if dtcode not in _DT:
        raise ValueError(f"{filename}: unsupported NIfTI datatype {dtcode}")
    dt = np.dtype(end + _DT[dtcode])
    pixdim = struct.unpack_from(end + "8f", raw, 76)
    vox_offset = int(struct.unpack_from(end + "f", raw, 108)[0])
    slope, inter = struct.unpack_from(end + "2f", raw, 112)
    qform, sform = struct.unpack_from(end + "2h", raw, 252)
    if sform > 0:
        affine = np.eye(4)
        for r, off in enumerate((280, 296, 312)):
            affine[r] = struct.unpack_from(end + "4f", raw, off)
    else:  # no sform: scale-only fallback from pixdim
        affine = np.diag([pixdim[1], pixdim[2], pixdim[3], 1.0])
    count = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=count, offset=vox_offset).reshape(shape, order="F")
    if np.isfinite(slope) and slope != 0 and not (slope == 1 and inter == 0):
        data = data * slope + (inter if np.isfinite(inter) else 0.0)
    descrip = raw[148:228].split(b"\x00", 1)[0]
    metadata: Dict[str, Any] = {
        "voxel_size": tuple(float(p) for p in pixdim[1:4]),
        "dimensions": shape,
        "affine_matrix": affine.tolist(),
        "data_type": str(dt.newbyteorder("=")),
        "description": str(descrip),  # str(bytes) like the reference's str(header['descrip'])
    }
    if raw[348] != 0:  # extensions present
        pos = HDR + 4
        while pos + 8 <= vox_offset:
            esize, ecode = struct.unpack_from(end + "2i", raw, pos)
            if esize < 8 or pos + esize > vox_offset:
                break
            if ecode == 44:
                try:
                    metadata.update(json.loads(raw[pos + 8:pos + esize].rstrip(b"\x00").decode("utf-8")))
                except Exception:
                    pass
            pos += esize
    return np.ascontiguousarray(data), metadata
