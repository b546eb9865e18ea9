"""Device-side engine: torch tensors for memory/streams, libpvdose (CUDA, sm_100a) for the arithmetic.

PyTorch is plumbing here (device memory, streams, pinned staging, torch.distributed); every
number is produced by the hand-written kernels behind the C ABI.  No CPU fallback: without a
CUDA device or without libpvdose.so the constructors raise.
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import ALGO_AUTO, BOUNDARY_REFERENCE, BOUNDARY_SAME, MAX_T, PvdoseError, get_lib

BOUNDARY_IDS = {"reference": BOUNDARY_REFERENCE, "circular": BOUNDARY_REFERENCE, "same": BOUNDARY_SAME, "zero": BOUNDARY_SAME}


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "pyvoxeldosimetry_b200 needs a CUDA device (B200, sm_100a): the kernel-convolution dose path has no CPU fallback"
        )
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if dev.type != "cuda":
        raise RuntimeError(f"device {dev} is not a CUDA device")
    return dev


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_NP_STAGE_DTYPES = {np.dtype(np.float32): _capi.DTYPE_F32, np.dtype(np.float64): _capi.DTYPE_F64,
                    np.dtype(np.int16): _capi.DTYPE_I16, np.dtype(np.uint16): _capi.DTYPE_U16}
STAGE_MIN_BYTES = 1 << 20  # smaller host arrays take the plain torch copy


class HostStager:
    """Pinned-ring + host-thread staging between pageable host ndarrays and device tensors (pvd_stager_*,
    csrc/host_stage.cuh).  One per device, created on first use."""

    _per_device: Dict[str, "HostStager"] = {}
    _lock = threading.Lock()

    @classmethod
    def get(cls, device: torch.device) -> "HostStager":
        key = str(device)
        with cls._lock:
            st = cls._per_device.get(key)
            if st is None:
                st = cls._per_device[key] = HostStager(device)
            return st

    def __init__(self, device: torch.device):
        self.device = device
        self.lib = get_lib()
        with torch.cuda.device(device):
            self.handle = self.lib.stager_create(0, 0, 0)

    def upload(self, arr: np.ndarray, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """C-contiguous float32 / float64 / int16 / uint16 ndarray -> device tensor (float64 narrowed to float32 on the
        host; uint16 arrives in an int16 tensor, bit for bit).  Enqueued on the current stream; `arr` is free on return."""
        code = _NP_STAGE_DTYPES[arr.dtype]
        tdt = torch.float32 if code in (_capi.DTYPE_F32, _capi.DTYPE_F64) else torch.int16
        if out is None:
            out = torch.empty(arr.shape, dtype=tdt, device=self.device)
        with torch.cuda.device(self.device):
            self.lib.stage_h2d(self.handle, arr.ctypes.data, code, out.data_ptr(), arr.size, _stream_ptr(self.device))
        return out

    def download(self, dev: torch.Tensor, out: np.ndarray) -> np.ndarray:
        """float32 device tensor -> C-contiguous float32 / float64 host ndarray (complete on return)."""
        code = _NP_STAGE_DTYPES[out.dtype]
        with torch.cuda.device(self.device):
            self.lib.stage_d2h(self.handle, dev.data_ptr(), out.ctypes.data, code, dev.numel(), _stream_ptr(self.device))
        return out


def _stageable(a: np.ndarray) -> bool:
    return a.dtype in _NP_STAGE_DTYPES and a.flags.c_contiguous and a.nbytes >= STAGE_MIN_BYTES


def to_device_f32(x, device: torch.device) -> torch.Tensor:
    """Host ndarray / tensor -> contiguous float32 CUDA tensor (no copy when already there).  Large float32 / float64
    ndarrays go through the staging engine (host threads + pinned ring; float64 is narrowed on the host)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype in (np.float32, np.float64) and _stageable(a):
            return HostStager.get(device).upload(a)
        if a.dtype != np.float32:
            a = a.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def to_host(dev: torch.Tensor, out=None, want64: bool = False) -> np.ndarray:
    """float32 device tensor -> host ndarray, complete on return.
      out=None        the result lives in pinned memory from torch's caching host allocator (one straight D2H at link
                      speed, no host copy); the ndarray owns the block, which returns to the cache when it is dropped;
                      want64 -> a float64 ndarray filled by the staging threads (the reference returns float64);
      out=ndarray     filled through the staging engine (float32 or float64, C-contiguous);
      out=torch CPU tensor (pinned or not): plain copy."""
    device = dev.device
    if out is None:
        if want64:
            res = np.empty(tuple(dev.shape), dtype=np.float64)
            if _stageable(res):
                return HostStager.get(device).download(dev, res)
            res[...] = dev.cpu().numpy()
            return res
        host = torch.empty(tuple(dev.shape), dtype=torch.float32, pin_memory=True)
        host.copy_(dev, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        return host.numpy()
    if isinstance(out, torch.Tensor):
        out.copy_(dev, non_blocking=False)
        return out.numpy()
    if tuple(out.shape) != tuple(dev.shape):
        raise ValueError("out has the wrong shape")
    if _stageable(out) and out.dtype in (np.float32, np.float64):
        return HostStager.get(device).download(dev, out)
    out[...] = dev.cpu().numpy()
    return out


class ConvPlan:
    """A reusable convolution plan: transform sizes, workspace, cached kernel spectrum.

    ``boundary='reference'`` reproduces core/kernel_convolution.py:71-74 (circular, origin-anchored
    kernel); ``boundary='same'`` is the zero-boundary, centred variant.  ``ex`` gives the expert
    geometry (slab decomposition): dict(m=, out_lo=, out_n=).
    """

    def __init__(self, shape: Sequence[int], kshape: Sequence[int], boundary: str = "reference", device=None,
                 algo: int = ALGO_AUTO, ex: Optional[dict] = None):
        self.device = require_cuda(device)
        self.lib = get_lib()
        self.shape = tuple(int(s) for s in shape)
        self.kshape = tuple(int(s) for s in kshape)
        if len(self.shape) != 3 or len(self.kshape) != 3:
            raise ValueError("activity and kernel must be 3-D")
        self.boundary = boundary
        with torch.cuda.device(self.device):
            if ex is not None:
                self.handle = self.lib.plan_create_ex(self.shape, ex.get("m", (0, 0, 0)), ex["out_lo"], ex["out_n"], self.kshape, algo)
            else:
                if boundary not in BOUNDARY_IDS:
                    raise ValueError(f"unknown boundary mode {boundary!r} (use 'reference' or 'same')")
                self.handle = self.lib.plan_create(self.shape, self.kshape, BOUNDARY_IDS[boundary], algo)
            self.info = self.lib.plan_info(self.handle)
            self.out_shape = tuple(self.info.out_n)
            self.fft_shape = tuple(self.info.m)
            nbytes = self.lib.plan_workspace_bytes(self.handle)
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.lib.plan_set_workspace(self.handle, self.workspace.data_ptr(), nbytes, _stream_ptr(self.device))
        self._kernel_dev: Optional[torch.Tensor] = None

    # -------------------------------------------------------------- kernel
    def set_kernel(self, kernel) -> None:
        k = to_device_f32(kernel, self.device)
        if tuple(k.shape) != self.kshape:
            raise ValueError(f"kernel shape {tuple(k.shape)} does not match the plan {self.kshape}")
        with torch.cuda.device(self.device):
            self.lib.plan_set_kernel(self.handle, k.data_ptr(), _stream_ptr(self.device))
        self._kernel_dev = k

    # -------------------------------------------------------------- execute (device resident)
    def execute(self, acts: Sequence[torch.Tensor], weights: Optional[Sequence[float]] = None,
                density: Optional[torch.Tensor] = None, rho_ref: float = 1.0, rho_min: float = 0.1,
                rho_cut: float = 0.0, scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """dose = scale * conv(sum_t w_t act_t, kernel) [* rho_ref / max(rho, rho_min)] - all tensors on device."""
        if len(acts) < 1:
            raise ValueError("No activity maps provided")
        for a in acts:
            if a.device != self.device or a.dtype != torch.float32 or not a.is_contiguous() or tuple(a.shape) != self.shape:
                raise ValueError("activity tensors must be contiguous float32 CUDA tensors of the plan shape")
        if density is not None and (density.device != self.device or density.dtype != torch.float32
                                    or not density.is_contiguous() or tuple(density.shape) != self.out_shape):
            raise ValueError("density must be a contiguous float32 CUDA tensor of the output shape")
        if out is None:
            out = torch.empty(self.out_shape, dtype=torch.float32, device=self.device)
        acts = list(acts)
        w = None if weights is None else [float(x) for x in weights]
        with torch.cuda.device(self.device):
            stream = _stream_ptr(self.device)
            if self.info.algo == _capi.ALGO_DIRECT and (len(acts) > 1 or w is not None):
                # the direct (TMA) kernel takes one volume: fold the time-weighted sum first (linearity)
                acts = [weighted_sum(acts, w if w is not None else [1.0] * len(acts))]
                w = None
            if len(acts) > MAX_T:  # fold the tail into one volume first (linearity)
                if w is None:
                    w = [1.0] * len(acts)
                acc = torch.empty(self.shape, dtype=torch.float32, device=self.device)
                first = True
                while len(acts) > MAX_T - 1:
                    chunk, cw = acts[: MAX_T - 1], w[: MAX_T - 1]
                    acts, w = acts[MAX_T - 1:], w[MAX_T - 1:]
                    ptrs, ws = [c.data_ptr() for c in chunk], list(cw)
                    if not first:
                        ptrs.append(acc.data_ptr())
                        ws.append(1.0)
                    self.lib.weighted_sum(ptrs, ws, acc.data_ptr(), acc.numel(), stream)
                    first = False
                acts, w = acts + [acc], w + [1.0]
            self.lib.conv_execute(self.handle, [a.data_ptr() for a in acts], w,
                                  None if density is None else density.data_ptr(),
                                  rho_ref, rho_min, rho_cut, scale, out.data_ptr(), stream)
        return out

    def execute_batch(self, acts: Sequence[Sequence[torch.Tensor]], weights: Optional[Sequence[float]] = None,
                      densities: Optional[Sequence[Optional[torch.Tensor]]] = None, rho_ref: float = 1.0, rho_min: float = 0.1,
                      rho_cut: float = 0.0, scale: float = 1.0, outs: Optional[Sequence[torch.Tensor]] = None) -> list:
        """`len(acts)` independent volume sets (acts[b] = the T time-point tensors of volume b, at most MAX_T) through ONE
        pvd_conv_execute_batch call; same per-volume semantics as execute().  FFT algorithm only."""
        B = len(acts)
        T = len(acts[0]) if B else 1
        if self.info.algo != _capi.ALGO_FFT or T > MAX_T or (B and T < 1):
            raise ValueError("execute_batch: FFT plans, 1..MAX_T time points per volume")
        for vol in acts:
            if len(vol) != T:
                raise ValueError("every volume of a batch needs the same number of time points")
            for a in vol:
                if a.device != self.device or a.dtype != torch.float32 or not a.is_contiguous() or tuple(a.shape) != self.shape:
                    raise ValueError("activity tensors must be contiguous float32 CUDA tensors of the plan shape")
        if densities is not None:
            if len(densities) != B:
                raise ValueError("one density entry (tensor or None) per volume")
            for d in densities:
                if d is not None and (d.device != self.device or d.dtype != torch.float32 or not d.is_contiguous()
                                      or tuple(d.shape) != self.out_shape):
                    raise ValueError("density must be a contiguous float32 CUDA tensor of the output shape")
        if outs is None:
            outs = [torch.empty(self.out_shape, dtype=torch.float32, device=self.device) for _ in range(B)]
        if len(outs) != B:
            raise ValueError("one output tensor per volume")
        with torch.cuda.device(self.device):
            self.lib.conv_execute_batch(self.handle, [[a.data_ptr() for a in vol] for vol in acts],
                                        None if weights is None else [float(x) for x in weights],
                                        None if densities is None else [None if d is None else d.data_ptr() for d in densities],
                                        rho_ref, rho_min, rho_cut, scale, [o.data_ptr() for o in outs], _stream_ptr(self.device))
        return list(outs)

    def check_device_errors(self) -> None:
        """Synchronise the current stream and raise PvdoseError if a device-side watchdog fired (a TMA tile copy that
        never completed); results of that execute are then invalid.  Tests and smoke() call it after executing."""
        with torch.cuda.device(self.device):
            self.lib.plan_check_device_errors(self.handle, _stream_ptr(self.device))

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.plan_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class PlanCache:
    """Plans keyed by (shape, kshape, boundary, device); the kernel spectrum is rebuilt only when the
    kernel content changes (the reference recomputes fftn(kernel) on every call)."""

    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self._plans: Dict[tuple, ConvPlan] = {}
        self._kernel_tag: Dict[tuple, object] = {}
        self._lock = threading.Lock()

    def get(self, shape, kshape, boundary, device, kernel_tag, kernel_provider, algo: int = ALGO_AUTO) -> ConvPlan:
        key = (tuple(shape), tuple(kshape), boundary, str(device), algo)
        with self._lock:
            plan = self._plans.get(key)
            if plan is None:
                while len(self._plans) >= self.capacity:
                    old_key = next(iter(self._plans))
                    self._plans.pop(old_key).close()
                    self._kernel_tag.pop(old_key, None)
                plan = ConvPlan(shape, kshape, boundary, device, algo)
                self._plans[key] = plan
            if self._kernel_tag.get(key) != kernel_tag:
                plan.set_kernel(kernel_provider())
                self._kernel_tag[key] = kernel_tag
            return plan

    def clear(self):
        with self._lock:
            for p in self._plans.values():
                p.close()
            self._plans.clear()
            self._kernel_tag.clear()


# ---------------------------------------------------------------------------- elementwise ops
def kernel_eval_radial(beta_terms, photon_terms, scaling: float, spacing: Sequence[float], grid: Sequence[int],
                       device=None) -> torch.Tensor:
    """Evaluate the radial dose-point-kernel model on the voxel grid, on the device (float32 tensor)."""
    dev = require_cuda(device)
    out = torch.empty(tuple(int(g) for g in grid), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        get_lib().kernel_eval_radial(beta_terms, photon_terms, scaling, spacing, grid, out.data_ptr(), _stream_ptr(dev))
    return out


def hu_to_density(hu: torch.Tensor, knots) -> torch.Tensor:
    dev = require_cuda(hu.device)
    if hu.dtype not in (torch.int16, torch.float32):
        hu = hu.to(torch.float32)
    hu = hu.contiguous()
    rho = torch.empty(hu.shape, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        get_lib().hu_to_density(hu.data_ptr(), hu.dtype == torch.int16, knots, rho.data_ptr(), hu.numel(), _stream_ptr(dev))
    return rho


def weighted_sum(vols: Sequence[torch.Tensor], weights: Sequence[float]) -> torch.Tensor:
    dev = require_cuda(vols[0].device)
    out = torch.empty_like(vols[0])
    lib = get_lib()
    with torch.cuda.device(dev):
        stream = _stream_ptr(dev)
        vols, weights = list(vols), [float(w) for w in weights]
        first = True
        while vols:
            take = MAX_T if first else MAX_T - 1
            chunk, cw = vols[:take], weights[:take]
            vols, weights = vols[take:], weights[take:]
            ptrs, ws = [c.data_ptr() for c in chunk], list(cw)
            if not first:
                ptrs.append(out.data_ptr())
                ws.append(1.0)
            lib.weighted_sum(ptrs, ws, out.data_ptr(), out.numel(), stream)
            first = False
    return out


MAX_J = 16


def weighted_combine(vols: Sequence[torch.Tensor], W) -> list:
    """[sum_t W[j][t] * vols[t] for j] with every volume read once per pass of <= MAX_J outputs (pvd_weighted_combine).
    More than MAX_T volumes: the weights are applied chunk by chunk, later chunks accumulate onto the outputs in place."""
    dev = require_cuda(vols[0].device)
    W = [[float(x) for x in row] for row in W]
    T = len(vols)
    if any(len(row) != T for row in W):
        raise ValueError("every weight row needs one entry per volume")
    for v in vols:
        if v.dtype != torch.float32 or v.device != vols[0].device or v.shape != vols[0].shape:
            raise ValueError("weighted_combine takes float32 CUDA tensors of one shape on one device")
    vols = [v if v.is_contiguous() else v.contiguous() for v in vols]
    outs = [torch.empty_like(vols[0]) for _ in W]
    lib = get_lib()
    with torch.cuda.device(dev):
        stream = _stream_ptr(dev)
        for j0 in range(0, len(W), MAX_J):
            rows, oj = W[j0:j0 + MAX_J], outs[j0:j0 + MAX_J]
            if T <= MAX_T:
                lib.weighted_combine([v.data_ptr() for v in vols], rows, [o.data_ptr() for o in oj], vols[0].numel(), stream)
                continue
            # long series: each output is (previous partial sum) + the next MAX_T - 1 volumes; J partial sums ride along
            # one at a time because a pass takes at most MAX_T inputs
            for o, row in zip(oj, rows):
                lib.weighted_combine([v.data_ptr() for v in vols[:MAX_T]], [row[:MAX_T]], [o.data_ptr()], o.numel(), stream)
                for t0 in range(MAX_T, T, MAX_T - 1):
                    chunk = vols[t0:t0 + MAX_T - 1]
                    lib.weighted_combine([v.data_ptr() for v in chunk] + [o.data_ptr()], [row[t0:t0 + MAX_T - 1] + [1.0]],
                                         [o.data_ptr()], o.numel(), stream)
    return outs


def monoexp_integral(A0: torch.Tensor, lam: torch.Tensor, t_limit: float) -> torch.Tensor:
    dev = require_cuda(A0.device)
    out = torch.empty_like(A0)
    with torch.cuda.device(dev):
        get_lib().monoexp_integral(A0.data_ptr(), lam.data_ptr(), float(t_limit), out.data_ptr(), A0.numel(), _stream_ptr(dev))
    return out


def density_scale(dose: torch.Tensor, density: torch.Tensor, rho_ref=1.0, rho_min=0.1, rho_cut=0.0, scale=1.0) -> torch.Tensor:
    dev = require_cuda(dose.device)
    out = torch.empty_like(dose)
    with torch.cuda.device(dev):
        get_lib().density_scale(dose.data_ptr(), density.data_ptr(), rho_ref, rho_min, rho_cut, scale, out.data_ptr(),
                                dose.numel(), _stream_ptr(dev))
    return out


# ---------------------------------------------------------------------------- steps either side of the convolution
def monoexp_fit(vols: Sequence[torch.Tensor], times: Sequence[float], weights: Optional[Sequence[float]], lambda0: float,
                t_limit: float, want_params: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """Per-voxel weighted mono-exponential fit + integral on the device.
    -> (params [2, *shape] or None, accumulated [*shape]); inputs are float32 CUDA tensors of one shape."""
    dev = require_cuda(vols[0].device)
    T = len(vols)
    if T > MAX_T:
        raise ValueError(f"at most {MAX_T} time points")
    shape = tuple(vols[0].shape)
    params = torch.empty((2,) + shape, dtype=torch.float32, device=dev) if want_params else None
    acc = torch.empty(shape, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        get_lib().monoexp_fit([v.data_ptr() for v in vols], times, weights, float(lambda0), float(t_limit),
                              params[0].data_ptr() if want_params else None, params[1].data_ptr() if want_params else None,
                              acc.data_ptr(), acc.numel(), _stream_ptr(dev))
    return params, acc


def ct_prepare(hu: torch.Tensor, metal_threshold: float, knots=None, ranges=None, want_corrected: bool = True):
    """One pass over a float32 HU volume -> (corrected HU | None, density | None, tissue bit labels | None)."""
    dev = require_cuda(hu.device)
    if hu.dim() != 3:
        raise ValueError("CT volume must be 3-D")
    hu = hu.to(torch.float32).contiguous()
    corrected = torch.empty_like(hu) if want_corrected else None
    rho = torch.empty_like(hu) if knots else None
    labels = torch.empty(hu.shape, dtype=torch.uint8, device=dev) if ranges else None
    with torch.cuda.device(dev):
        get_lib().ct_prepare(hu.data_ptr(), tuple(hu.shape), metal_threshold, knots, ranges,
                             None if corrected is None else corrected.data_ptr(), None if rho is None else rho.data_ptr(),
                             None if labels is None else labels.data_ptr(), _stream_ptr(dev))
    return corrected, rho, labels


def roi_minmax(dose: torch.Tensor, mask: torch.Tensor) -> Tuple[float, float, int]:
    dev = require_cuda(dose.device)
    scratch = torch.empty(16, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        return get_lib().roi_minmax(dose.data_ptr(), mask.data_ptr(), mask.dtype == torch.float32, dose.numel(),
                                    scratch.data_ptr(), _stream_ptr(dev))


def dvh_histogram(dose: torch.Tensor, mask: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """Counts per bin (int64 CUDA tensor) of dose[mask > 0] for the uniform float32 edge array `edges` (CUDA)."""
    dev = require_cuda(dose.device)
    bins = edges.numel() - 1
    hist = torch.empty(bins, dtype=torch.int64, device=dev)
    e = edges.cpu()
    with torch.cuda.device(dev):
        get_lib().dvh_histogram(dose.data_ptr(), mask.data_ptr(), mask.dtype == torch.float32, dose.numel(), edges.data_ptr(),
                                bins, float(e[0]), float(e[-1]), hist.data_ptr(), _stream_ptr(dev))
    return hist
