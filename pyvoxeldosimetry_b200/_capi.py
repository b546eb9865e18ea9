"""ctypes binding of libpvdose.so (include/pvdose.h).

The binding is pointer-agnostic: every array argument is an integer address.  The product hands
it ``torch.Tensor.data_ptr()`` of CUDA tensors; there is NO CPU fallback - if the shared library
is missing the import of the engine fails loudly (``PvdoseLibraryError``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libpvdose.so")

PVD_OK = 0
ERR_NAMES = {-1: "PVD_ERR_INVALID", -2: "PVD_ERR_CUDA", -3: "PVD_ERR_STATE", -4: "PVD_ERR_NONFINITE", -5: "PVD_ERR_UNSUPPORTED"}
BOUNDARY_REFERENCE, BOUNDARY_SAME = 0, 1
ALGO_AUTO, ALGO_FFT, ALGO_DIRECT, ALGO_FFT_UNPIPELINED = 0, 1, 2, 3
MAX_T = 16

# every symbol include/pvdose.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "pvd_version", "pvd_build_id", "pvd_last_error", "pvd_good_fft_size", "pvd_good_fft_size_axis", "pvd_plan_create", "pvd_plan_create_ex", "pvd_plan_get_info",
    "pvd_plan_workspace_bytes", "pvd_plan_set_workspace", "pvd_plan_set_kernel", "pvd_conv_execute", "pvd_conv_execute_batch", "pvd_conv_forward_planes", "pvd_conv_finish", "pvd_conv_middle", "pvd_conv_output_planes", "pvd_plan_reserve_sms", "pvd_stream_write_flag", "pvd_stream_wait_flag_geq", "pvd_copy_async", "pvd_plan_destroy", "pvd_plan_set_profiling", "pvd_plan_get_pass_times", "pvd_plan_check_device_errors",
    "pvd_kernel_eval_radial", "pvd_hu_to_density_f32", "pvd_hu_to_density_i16", "pvd_weighted_sum", "pvd_weighted_combine", "pvd_monoexp_integral",
    "pvd_density_scale", "pvd_monoexp_fit", "pvd_ct_prepare", "pvd_roi_minmax", "pvd_dvh_histogram",
    "pvd_stager_create", "pvd_stager_destroy", "pvd_stage_h2d", "pvd_stage_d2h", "pvd_i16_to_f32",
]
DTYPE_F32, DTYPE_F64, DTYPE_I16, DTYPE_U16 = 0, 1, 2, 3


class PvdoseLibraryError(RuntimeError):
    pass


class PvdoseError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class PlanInfo(C.Structure):
    _fields_ = [
        ("n", C.c_int * 3), ("m", C.c_int * 3), ("out_lo", C.c_int * 3), ("out_n", C.c_int * 3), ("k", C.c_int * 3),
        ("algo", C.c_int), ("passes", C.c_int), ("workspace_bytes", C.c_size_t), ("hbm_bytes_per_execute", C.c_double),
    ]


class RadialModel(C.Structure):
    _fields_ = [("n_beta", C.c_int), ("n_photon", C.c_int), ("beta_range", C.c_double * 4), ("beta_amp", C.c_double * 4),
                ("phot_mu", C.c_double * 4), ("phot_amp", C.c_double * 4), ("scaling", C.c_double)]


def _i3(v: Sequence[int]):
    return (C.c_int * 3)(*[int(x) for x in v])


class PvdLib:
    """Thin, typed view of the shared library."""

    def __init__(self, path: Optional[str] = None):
        path = path or os.environ.get("PVDOSE_LIB", DEFAULT_LIB)
        if not os.path.exists(path):
            raise PvdoseLibraryError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback for the dose path."
            )
        try:
            self.dll = C.CDLL(path)
        except OSError as e:  # pragma: no cover
            raise PvdoseLibraryError(f"cannot load {path}: {e}") from e
        self.path = path
        d = self.dll
        vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
        d.pvd_version.restype = C.c_int
        d.pvd_last_error.restype = C.c_char_p
        d.pvd_build_id.restype = C.c_char_p
        d.pvd_good_fft_size.argtypes = [C.c_int]
        d.pvd_good_fft_size_axis.argtypes = [C.c_int, C.c_int]
        d.pvd_plan_create.argtypes = [C.POINTER(vp), ip, ip, C.c_int, C.c_int]
        d.pvd_plan_create_ex.argtypes = [C.POINTER(vp), ip, ip, ip, ip, ip, C.c_int]
        d.pvd_plan_get_info.argtypes = [vp, C.POINTER(PlanInfo)]
        d.pvd_plan_workspace_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
        d.pvd_plan_set_workspace.argtypes = [vp, vp, C.c_size_t, vp]
        d.pvd_plan_set_kernel.argtypes = [vp, vp, vp]
        d.pvd_conv_execute.argtypes = [vp, C.POINTER(vp), fp, C.c_int, vp, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp]
        d.pvd_conv_execute_batch.argtypes = [vp, C.POINTER(vp), fp, C.c_int, C.POINTER(vp), C.c_float, C.c_float, C.c_float, C.c_float,
                                             C.POINTER(vp), C.c_int, vp]
        d.pvd_conv_forward_planes.argtypes = [vp, C.POINTER(vp), fp, C.c_int, C.c_float, C.c_int, C.c_int, vp]
        d.pvd_conv_finish.argtypes = [vp, vp, C.c_float, C.c_float, vp, vp]
        d.pvd_conv_middle.argtypes = [vp, vp]
        d.pvd_plan_reserve_sms.argtypes = [vp, C.c_int]
        d.pvd_stream_write_flag.argtypes = [vp, C.c_uint32, vp]
        d.pvd_stream_wait_flag_geq.argtypes = [vp, C.c_uint32, vp]
        d.pvd_copy_async.argtypes = [vp, vp, C.c_size_t, vp]
        d.pvd_conv_output_planes.argtypes = [vp, vp, C.c_float, C.c_float, vp, C.c_int, C.c_int, vp]
        d.pvd_plan_destroy.argtypes = [vp]
        d.pvd_plan_set_profiling.argtypes = [vp, C.c_int]
        d.pvd_plan_get_pass_times.argtypes = [vp, fp, C.POINTER(C.c_double), C.POINTER(C.c_char_p), C.c_int]
        d.pvd_plan_check_device_errors.argtypes = [vp, vp]
        d.pvd_kernel_eval_radial.argtypes = [C.POINTER(RadialModel), C.POINTER(C.c_double), ip, vp, vp]
        d.pvd_hu_to_density_f32.argtypes = [vp, fp, C.c_int, vp, C.c_size_t, vp]
        d.pvd_hu_to_density_i16.argtypes = [vp, fp, C.c_int, vp, C.c_size_t, vp]
        d.pvd_weighted_sum.argtypes = [C.POINTER(vp), fp, C.c_int, vp, C.c_size_t, vp]
        d.pvd_weighted_combine.argtypes = [C.POINTER(vp), C.c_int, fp, C.POINTER(vp), C.c_int, C.c_size_t, vp]
        d.pvd_monoexp_integral.argtypes = [vp, vp, C.c_float, vp, C.c_size_t, vp]
        d.pvd_density_scale.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, vp, C.c_size_t, vp]
        d.pvd_monoexp_fit.argtypes = [C.POINTER(vp), fp, fp, C.c_int, C.c_float, C.c_float, vp, vp, vp, C.c_size_t, vp]
        d.pvd_ct_prepare.argtypes = [vp, ip, C.c_float, fp, C.c_int, fp, C.c_int, vp, vp, vp, vp]
        d.pvd_roi_minmax.argtypes = [vp, vp, C.c_int, C.c_size_t, vp, fp, fp, C.POINTER(C.c_ulonglong), vp]
        d.pvd_dvh_histogram.argtypes = [vp, vp, C.c_int, C.c_size_t, vp, C.c_int, C.c_float, C.c_float, vp, vp]
        d.pvd_stager_create.argtypes = [C.POINTER(vp), C.c_int, C.c_size_t, C.c_int]
        d.pvd_stager_destroy.argtypes = [vp]
        d.pvd_stage_h2d.argtypes = [vp, vp, C.c_int, vp, C.c_size_t, vp]
        d.pvd_stage_d2h.argtypes = [vp, vp, vp, C.c_int, C.c_size_t, vp]
        d.pvd_i16_to_f32.argtypes = [vp, C.c_int, C.c_float, C.c_float, vp, C.c_size_t, vp]

    # ------------------------------------------------------------------ helpers
    def check(self, rc: int):
        if rc != PVD_OK:
            raise PvdoseError(rc, (self.dll.pvd_last_error() or b"").decode())

    def version(self) -> int:
        return self.dll.pvd_version()

    def build_id(self) -> str:
        return (self.dll.pvd_build_id() or b"").decode()

    def good_fft_size(self, n: int, axis: int = 0) -> int:
        return self.dll.pvd_good_fft_size_axis(int(n), int(axis))

    def plan_create(self, n, k, boundary: int, algo: int = ALGO_AUTO) -> int:
        h = C.c_void_p()
        self.check(self.dll.pvd_plan_create(C.byref(h), _i3(n), _i3(k), boundary, algo))
        return h.value

    def plan_create_ex(self, n, m, out_lo, out_n, k, algo: int = ALGO_AUTO) -> int:
        h = C.c_void_p()
        self.check(self.dll.pvd_plan_create_ex(C.byref(h), _i3(n), _i3(m), _i3(out_lo), _i3(out_n), _i3(k), algo))
        return h.value

    def plan_info(self, plan: int) -> PlanInfo:
        info = PlanInfo()
        self.check(self.dll.pvd_plan_get_info(plan, C.byref(info)))
        return info

    def plan_workspace_bytes(self, plan: int) -> int:
        b = C.c_size_t()
        self.check(self.dll.pvd_plan_workspace_bytes(plan, C.byref(b)))
        return b.value

    def plan_set_workspace(self, plan: int, ptr: int, nbytes: int, stream: int = 0):
        self.check(self.dll.pvd_plan_set_workspace(plan, ptr, nbytes, stream))

    def plan_set_kernel(self, plan: int, kernel_ptr: int, stream: int = 0):
        self.check(self.dll.pvd_plan_set_kernel(plan, kernel_ptr, stream))

    def conv_execute(self, plan: int, act_ptrs: Sequence[int], weights: Optional[Sequence[float]], density_ptr: Optional[int],
                     rho_ref: float, rho_min: float, rho_cut: float, scale: float, dose_ptr: int, stream: int = 0):
        T = len(act_ptrs)
        ptrs = (C.c_void_p * T)(*act_ptrs)
        w = (C.c_float * T)(*[float(x) for x in weights]) if weights is not None else None
        self.check(self.dll.pvd_conv_execute(plan, ptrs, w, T, density_ptr, rho_ref, rho_min, rho_cut, scale, dose_ptr, stream))

    def conv_execute_batch(self, plan: int, act_ptrs: Sequence[Sequence[int]], weights: Optional[Sequence[float]],
                           density_ptrs: Optional[Sequence[Optional[int]]], rho_ref: float, rho_min: float, rho_cut: float, scale: float,
                           dose_ptrs: Sequence[int], stream: int = 0):
        """act_ptrs[b][t]; density_ptrs None or one entry (pointer or None) per volume; dose_ptrs one per volume."""
        B = len(act_ptrs)
        T = len(act_ptrs[0]) if B else 1
        if any(len(a) != T for a in act_ptrs) or len(dose_ptrs) != B or (density_ptrs is not None and len(density_ptrs) != B):
            raise ValueError("every volume of a batch needs T activity pointers, one output and (optionally) one density pointer")
        flat = (C.c_void_p * max(1, B * T))(*[p for a in act_ptrs for p in a])
        w = (C.c_float * T)(*[float(x) for x in weights]) if weights is not None else None
        den = (C.c_void_p * max(1, B))(*density_ptrs) if density_ptrs is not None else None
        outs = (C.c_void_p * max(1, B))(*dose_ptrs)
        self.check(self.dll.pvd_conv_execute_batch(plan, flat, w, T, den, rho_ref, rho_min, rho_cut, scale, outs, B, stream))

    def conv_forward_planes(self, plan: int, act_ptrs: Sequence[int], weights: Optional[Sequence[float]], gain: float, lo: int, hi: int,
                            stream: int = 0):
        T = len(act_ptrs)
        ptrs = (C.c_void_p * T)(*act_ptrs)
        w = (C.c_float * T)(*[float(x) for x in weights]) if weights is not None else None
        self.check(self.dll.pvd_conv_forward_planes(plan, ptrs, w, T, gain, lo, hi, stream))

    def conv_finish(self, plan: int, density_ptr: Optional[int], rho_min: float, rho_cut: float, dose_ptr: int, stream: int = 0):
        self.check(self.dll.pvd_conv_finish(plan, density_ptr, rho_min, rho_cut, dose_ptr, stream))

    def plan_reserve_sms(self, plan: int, n_sms: int):
        self.check(self.dll.pvd_plan_reserve_sms(plan, int(n_sms)))

    def stream_write_flag(self, flag_ptr: int, value: int, stream: int = 0):
        self.check(self.dll.pvd_stream_write_flag(flag_ptr, value, stream))

    def stream_wait_flag_geq(self, flag_ptr: int, value: int, stream: int = 0):
        self.check(self.dll.pvd_stream_wait_flag_geq(flag_ptr, value, stream))

    def copy_async(self, dst_ptr: int, src_ptr: int, nbytes: int, stream: int = 0):
        self.check(self.dll.pvd_copy_async(dst_ptr, src_ptr, nbytes, stream))

    def conv_middle(self, plan: int, stream: int = 0):
        self.check(self.dll.pvd_conv_middle(plan, stream))

    def conv_output_planes(self, plan: int, density_ptr: Optional[int], rho_min: float, rho_cut: float, dose_ptr: int, lo: int, hi: int,
                           stream: int = 0):
        self.check(self.dll.pvd_conv_output_planes(plan, density_ptr, rho_min, rho_cut, dose_ptr, lo, hi, stream))

    def plan_destroy(self, plan: int):
        self.dll.pvd_plan_destroy(plan)

    def plan_set_profiling(self, plan: int, enable: bool):
        self.check(self.dll.pvd_plan_set_profiling(plan, 1 if enable else 0))

    def plan_get_pass_times(self, plan: int):
        """-> [(name, ms, hbm_bytes)] of the last execute (synchronises its events)."""
        cap = 8
        ms, by, names = (C.c_float * cap)(), (C.c_double * cap)(), (C.c_char_p * cap)()
        n = self.dll.pvd_plan_get_pass_times(plan, ms, by, names, cap)
        if n < 0:
            self.check(n)
        return [(names[i].decode(), float(ms[i]), float(by[i])) for i in range(n)]

    def plan_check_device_errors(self, plan: int, stream: int = 0):
        """Synchronise `stream` and raise if a device-side TMA watchdog fired (see include/pvdose.h)."""
        self.check(self.dll.pvd_plan_check_device_errors(plan, stream))

    def kernel_eval_radial(self, beta_terms, photon_terms, scaling: float, spacing, grid, out_ptr: int, stream: int = 0):
        """beta_terms: [(range_mm, amplitude)], photon_terms: [(mu_per_cm, amplitude)]."""
        m = RadialModel()
        m.n_beta, m.n_photon, m.scaling = len(beta_terms), len(photon_terms), float(scaling)
        if m.n_beta > 4 or m.n_photon > 4:
            raise ValueError("at most 4 beta and 4 photon terms")
        for i, (r, a) in enumerate(beta_terms):
            m.beta_range[i], m.beta_amp[i] = float(r), float(a)
        for i, (mu, a) in enumerate(photon_terms):
            m.phot_mu[i], m.phot_amp[i] = float(mu), float(a)
        sp = (C.c_double * 3)(*[float(s) for s in spacing])
        self.check(self.dll.pvd_kernel_eval_radial(C.byref(m), sp, _i3(grid), out_ptr, stream))

    def hu_to_density(self, hu_ptr: int, is_i16: bool, knots, rho_ptr: int, n: int, stream: int = 0):
        flat = [float(v) for pair in knots for v in pair]
        arr = (C.c_float * len(flat))(*flat)
        fn = self.dll.pvd_hu_to_density_i16 if is_i16 else self.dll.pvd_hu_to_density_f32
        self.check(fn(hu_ptr, arr, len(flat) // 2, rho_ptr, n, stream))

    def weighted_sum(self, vol_ptrs: Sequence[int], weights: Sequence[float], out_ptr: int, n: int, stream: int = 0):
        T = len(vol_ptrs)
        ptrs = (C.c_void_p * T)(*vol_ptrs)
        w = (C.c_float * T)(*[float(x) for x in weights])
        self.check(self.dll.pvd_weighted_sum(ptrs, w, T, out_ptr, n, stream))

    def weighted_combine(self, vol_ptrs: Sequence[int], W: Sequence[Sequence[float]], out_ptrs: Sequence[int], n: int, stream: int = 0):
        """out[j] = sum_t W[j][t] * vol[t] in one pass (pvd_weighted_combine)."""
        T, J = len(vol_ptrs), len(out_ptrs)
        if len(W) != J or any(len(row) != T for row in W):
            raise ValueError("W must have one row of T weights per output")
        ptrs = (C.c_void_p * T)(*vol_ptrs)
        outs = (C.c_void_p * J)(*out_ptrs)
        w = (C.c_float * (J * T))(*[float(x) for row in W for x in row])
        self.check(self.dll.pvd_weighted_combine(ptrs, T, w, outs, J, n, stream))

    def monoexp_integral(self, a0_ptr: int, lam_ptr: int, t_limit: float, out_ptr: int, n: int, stream: int = 0):
        self.check(self.dll.pvd_monoexp_integral(a0_ptr, lam_ptr, t_limit, out_ptr, n, stream))

    def density_scale(self, dose_ptr: int, den_ptr: int, rho_ref: float, rho_min: float, rho_cut: float, scale: float,
                      out_ptr: int, n: int, stream: int = 0):
        self.check(self.dll.pvd_density_scale(dose_ptr, den_ptr, rho_ref, rho_min, rho_cut, scale, out_ptr, n, stream))


    def monoexp_fit(self, vol_ptrs: Sequence[int], times: Sequence[float], weights: Optional[Sequence[float]], lambda0: float,
                    t_limit: float, a0_ptr: Optional[int], lam_ptr: Optional[int], acc_ptr: Optional[int], n: int, stream: int = 0):
        T = len(vol_ptrs)
        ptrs = (C.c_void_p * T)(*vol_ptrs)
        t = (C.c_float * T)(*[float(x) for x in times])
        w = (C.c_float * T)(*[float(x) for x in weights]) if weights is not None else None
        self.check(self.dll.pvd_monoexp_fit(ptrs, t, w, T, lambda0, t_limit, a0_ptr, lam_ptr, acc_ptr, n, stream))

    def ct_prepare(self, hu_ptr: int, shape, metal_threshold: float, knots, ranges, corrected_ptr: Optional[int],
                   rho_ptr: Optional[int], labels_ptr: Optional[int], stream: int = 0):
        kflat = [float(v) for pair in (knots or []) for v in pair]
        rflat = [float(v) for pair in (ranges or []) for v in pair]
        karr = (C.c_float * len(kflat))(*kflat) if kflat else None
        rarr = (C.c_float * len(rflat))(*rflat) if rflat else None
        self.check(self.dll.pvd_ct_prepare(hu_ptr, _i3(shape), float(metal_threshold), karr, len(kflat) // 2, rarr, len(rflat) // 2,
                                           corrected_ptr, rho_ptr, labels_ptr, stream))

    def roi_minmax(self, dose_ptr: int, mask_ptr: int, mask_is_f32: bool, n: int, scratch_ptr: int, stream: int = 0):
        mn, mx, cnt = C.c_float(), C.c_float(), C.c_ulonglong()
        self.check(self.dll.pvd_roi_minmax(dose_ptr, mask_ptr, 1 if mask_is_f32 else 0, n, scratch_ptr, C.byref(mn), C.byref(mx),
                                           C.byref(cnt), stream))
        return mn.value, mx.value, cnt.value

    def dvh_histogram(self, dose_ptr: int, mask_ptr: int, mask_is_f32: bool, n: int, edges_ptr: int, bins: int, first: float,
                      last: float, hist_ptr: int, stream: int = 0):
        self.check(self.dll.pvd_dvh_histogram(dose_ptr, mask_ptr, 1 if mask_is_f32 else 0, n, edges_ptr, bins, first, last, hist_ptr, stream))


    # ------------------------------------------------------------------ host-buffer staging
    def stager_create(self, threads: int = 0, chunk_bytes: int = 0, ring_chunks: int = 0) -> int:
        h = C.c_void_p()
        self.check(self.dll.pvd_stager_create(C.byref(h), threads, chunk_bytes, ring_chunks))
        return h.value

    def stager_destroy(self, stager: int):
        self.dll.pvd_stager_destroy(stager)

    def stage_h2d(self, stager: int, host_ptr: int, dtype: int, dev_ptr: int, n: int, stream: int = 0):
        self.check(self.dll.pvd_stage_h2d(stager, host_ptr, dtype, dev_ptr, n, stream))

    def stage_d2h(self, stager: int, dev_ptr: int, host_ptr: int, dtype: int, n: int, stream: int = 0):
        self.check(self.dll.pvd_stage_d2h(stager, dev_ptr, host_ptr, dtype, n, stream))

    def i16_to_f32(self, in_ptr: int, is_unsigned: bool, slope: float, intercept: float, out_ptr: int, n: int, stream: int = 0):
        self.check(self.dll.pvd_i16_to_f32(in_ptr, 1 if is_unsigned else 0, slope, intercept, out_ptr, n, stream))


_LIB: Optional[PvdLib] = None


def get_lib() -> PvdLib:
    """The process-wide product library (libpvdose.so next to this file)."""
    global _LIB
    if _LIB is None:
        _LIB = PvdLib()
    return _LIB
