"""CPU SIMT-emulation harness (TEST INFRASTRUCTURE ONLY).

The CUDA sources of libpvdose also compile with g++ under -DPVD_EMULATE (see
pyvoxeldosimetry_b200/csrc/pvd_common.cuh): one std::thread per CUDA thread, a std::barrier for
__syncthreads.  That lets the `-m "not gpu"` suite drive the *real kernel code* (index math,
radix schedule, crop/pad logic, fused epilogues) through the real C ABI, with NumPy arrays standing
in for device memory.  The product never loads this library.
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
EMU_DIR = os.path.join(REPO, "tests", "_emu")
EMU_LIB = os.path.join(EMU_DIR, "libpvdose_emu.so")
CSRC = os.path.join(REPO, "pyvoxeldosimetry_b200", "csrc")


def build_emu(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(REPO, "include", "pvdose.h"))
    if not force and os.path.exists(EMU_LIB) and all(os.path.getmtime(EMU_LIB) >= os.path.getmtime(s) for s in srcs):
        return EMU_LIB
    os.makedirs(EMU_DIR, exist_ok=True)
    cus = [s for s in srcs if s.endswith(".cu")]
    cmd = ["g++", "-std=c++20", "-O2", "-x", "c++", "-DPVD_EMULATE", "-shared", "-fPIC", "-pthread", "-o", EMU_LIB] + cus
    subprocess.run(cmd, check=True)
    return EMU_LIB


def emu_lib():
    from pyvoxeldosimetry_b200._capi import PvdLib

    return PvdLib(build_emu())


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class EmuConv:
    """NumPy-memory twin of pyvoxeldosimetry_b200.engine.ConvPlan for the emulated library."""

    def __init__(self, lib, n, k, boundary=0, ex=None, algo=0):
        self.lib = lib
        if ex is None:
            self.plan = lib.plan_create(n, k, boundary, algo)
        else:
            self.plan = lib.plan_create_ex(n, ex["m"], ex["out_lo"], ex["out_n"], k, algo)
        self.info = lib.plan_info(self.plan)
        nbytes = lib.plan_workspace_bytes(self.plan)
        raw = np.zeros(nbytes + 256, dtype=np.uint8)
        off = (-raw.ctypes.data) % 256
        self.ws = raw[off : off + nbytes]
        lib.plan_set_workspace(self.plan, _ptr(self.ws), nbytes)
        self.out_shape = tuple(self.info.out_n)

    def set_kernel(self, kernel):
        self.kernel = np.ascontiguousarray(kernel, dtype=np.float32)
        self.lib.plan_set_kernel(self.plan, _ptr(self.kernel))

    def execute(self, acts, weights=None, density=None, rho_ref=1.0, rho_min=0.1, rho_cut=0.0, scale=1.0):
        acts = [np.ascontiguousarray(a, dtype=np.float32) for a in acts]
        den = None if density is None else np.ascontiguousarray(density, dtype=np.float32)
        out = np.full(self.out_shape, np.nan, dtype=np.float32)
        self.lib.conv_execute(self.plan, [_ptr(a) for a in acts], weights, None if den is None else _ptr(den),
                              rho_ref, rho_min, rho_cut, scale, _ptr(out))
        return out

    def close(self):
        self.lib.plan_destroy(self.plan)
