"""KernelConvolutionCalculator - drop-in for the reference class of the same name
(core/kernel_convolution.py:26-115) with the arithmetic on the GPU.

Reference semantics kept by default:
  * calculate_dose_rate == np.fft.ifftn(fftn(a) * fftn(kernel, a.shape)).real (kernel_convolution.py:71-74):
    circular over the activity grid, kernel anchored at the origin  -> ``boundary='reference'``;
  * calculate_absorbed_dose == trapezoid of the per-timepoint dose rates, hours -> seconds
    (kernel_convolution.py:94-106), evaluated as ONE convolution of sum_i w_i a_i (linearity);
  * the kernel is the factory's 64^3 grid at ``kernel_resolution`` (kernel_convolution.py:39-46).
Config keys (all optional): boundary ('reference'|'same'), kernel_grid, device, output_dtype
('float32'|'float64'), rho_ref, rho_min, rho_cut, scale, strict_reference, algo ('auto'|'fft'|'direct').
Repairs over the reference are listed in SURVEY.md section 8b; `strict_reference=True` turns the
behavioural ones off (the resample stub then raises instead of crashing with AttributeError).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import engine
from ..data.dose_kernels.kernel_factory import KernelFactory
from .dosimetry_base import DosimetryCalculator

HOURS_TO_SECONDS = 3600.0  # kernel_convolution.py:102


def trapezoid_weights(time_points: Sequence[float], unit_factor: float = 1.0) -> List[float]:
    """Weights w_i with sum_i w_i f(t_i) == the reference's trapezoid loops
    (kernel_convolution.py:101-104 with unit_factor=3600; activity_sampler.py:74-78 with 1)."""
    t = [float(x) for x in time_points]
    w = [0.0] * len(t)
    for i in range(len(t) - 1):
        d = (t[i + 1] - t[i]) * unit_factor
        w[i] += d / 2
        w[i + 1] += d / 2
    return w


def _same_spacing(voxel_size, res: float) -> bool:
    v = np.asarray(voxel_size, dtype=np.float64).reshape(-1)
    return v.size == 3 and bool(np.all(np.abs(v - res) <= 1e-9 * max(1.0, abs(res))))


class KernelConvolutionCalculator(DosimetryCalculator):
    def __init__(self, radionuclide: str, tissue_name: str, kernel_resolution: float = 1.0,
                 config: Optional[Dict[str, Any]] = None):
        super().__init__(radionuclide, tissue_name, config)
        self.device = engine.require_cuda(self.config.get("device"))
        self.kernel_factory = KernelFactory()
        self.kernel_resolution = kernel_resolution
        self.tissue_name = tissue_name
        self.boundary = self.config.get("boundary", "reference")
        if self.boundary not in engine.BOUNDARY_IDS:
            raise ValueError(f"unknown boundary mode {self.boundary!r} (use 'reference' or 'same')")
        self.kernel_grid = tuple(self.config.get("kernel_grid", (64, 64, 64)))  # kernel_convolution.py:45
        self.strict_reference = bool(self.config.get("strict_reference", False))
        # 'auto' picks the direct TMA-tiled convolution for small kernels in 'same' mode, the FFT path otherwise
        self.algo = {"auto": 0, "fft": 1, "direct": 2}[str(self.config.get("algo", "auto"))]
        self._plans = engine.PlanCache(capacity=int(self.config.get("plan_cache", 4)))
        self._kernel_version = 0
        self._kernel_host: Optional[np.ndarray] = None
        self._kernel_dev: Optional[torch.Tensor] = None
        self._aniso_kernels: Dict[tuple, torch.Tensor] = {}
        self._kernel_user = False
        self._load_dose_kernel()

    # ------------------------------------------------------------------ kernel state
    def _load_dose_kernel(self) -> None:
        self._kernel_dev = self.kernel_factory.get_kernel_device(
            nuclide=self.radionuclide, tissue_type=self.tissue_name, voxel_size=self.kernel_resolution,
            grid_size=self.kernel_grid, device=self.device)
        self._kernel_host = None
        self._kernel_user = False
        self._aniso_kernels.clear()
        self._kernel_version += 1

    @property
    def kernel(self) -> np.ndarray:
        """Public, assignable kernel (the reference exposes ``self.kernel`` as mutable state).  The returned array is a
        read-only host copy: in-place edits would never reach the device, so they raise instead of being lost -
        assign a new array (``calc.kernel = k``) to change the kernel."""
        if self._kernel_host is None:
            host = self._kernel_dev.cpu().numpy().astype(np.float64)
            host.setflags(write=False)
            self._kernel_host = host
        return self._kernel_host

    @kernel.setter
    def kernel(self, value) -> None:
        arr = np.asarray(value)
        if arr.ndim != 3:
            raise ValueError("kernel must be a 3-D array")
        host = np.array(arr, dtype=np.float64)  # private copy: later edits of the caller's array do not alias the device kernel
        host.setflags(write=False)
        self._kernel_host = host
        self._kernel_dev = engine.to_device_f32(arr, self.device)
        self._kernel_user = True   # a caller-supplied kernel is taken as sampled on the image grid, whatever voxel_size says
        self._aniso_kernels.clear()
        self._kernel_version += 1  # cached spectra are rebuilt on next use

    def _kernel_for(self, voxel_size) -> Tuple[torch.Tensor, object]:
        if voxel_size is None or self._kernel_user or _same_spacing(voxel_size, self.kernel_resolution):
            return self._kernel_dev, ("k", self._kernel_version)
        if len(tuple(voxel_size)) != 3:
            raise ValueError("voxel_size must be a tuple of length 3.")
        if self.strict_reference:
            raise NotImplementedError(
                "voxel_size differs from kernel_resolution: the reference's _resample_activity is a stub "
                "(core/kernel_convolution.py:108-115)")
        # A10: evaluate the dose voxel kernel on the image grid instead of resampling the activity
        sp = tuple(float(v) for v in voxel_size)
        if sp not in self._aniso_kernels:
            self._aniso_kernels[sp] = self.kernel_factory.get_kernel_device(
                self.radionuclide, self.tissue_name, sp, self.kernel_grid, device=self.device)
        return self._aniso_kernels[sp], ("sp", sp)

    # ------------------------------------------------------------------ core
    def _density_from_ct(self, ct_hu, shape) -> torch.Tensor:
        """CT volume in Hounsfield units (int16 as scanners store it, or float) -> device density map through the
        piecewise-linear HU table (`hu_knots` config, default tissue.HU_KNOTS); the conversion runs on the device,
        so an int16 CT crosses the PCIe link at 2 bytes per voxel."""
        from ..tissue.density import HU_KNOTS

        if tuple(ct_hu.shape) != tuple(shape):
            raise ValueError("ct_hu must have the shape of the activity map")
        if isinstance(ct_hu, torch.Tensor):
            t = ct_hu if ct_hu.dtype in (torch.int16, torch.float32) else ct_hu.to(torch.float32)
            t = t.to(self.device, non_blocking=True).contiguous()
        else:
            a = np.ascontiguousarray(ct_hu)
            if a.dtype not in (np.int16, np.float32, np.float64):
                a = a.astype(np.float32)
            if engine._stageable(a):
                t = engine.HostStager.get(self.device).upload(a)  # int16 stays int16 (2 bytes per voxel over the link)
            else:
                t = torch.from_numpy(a if a.dtype != np.float64 else a.astype(np.float32)).to(self.device, non_blocking=True)
        return engine.hu_to_density(t, self.config.get("hu_knots", HU_KNOTS))

    def _activity_to_device(self, m, rescale=None) -> torch.Tensor:
        """Activity volume -> float32 device tensor.  16-bit stored activity (the PET DICOM pixel data the reference
        rescales on the host, io/dicom.py:27-47) crosses the link at 2 bytes per voxel and becomes
        slope * stored + intercept on the device; `rescale` = (slope, intercept), default (1, 0)."""
        is16 = (isinstance(m, np.ndarray) and m.dtype in (np.int16, np.uint16)) or \
               (isinstance(m, torch.Tensor) and m.dtype in (torch.int16, torch.uint16))
        if not is16:
            if rescale is not None:
                raise ValueError("rescale=(slope, intercept) applies to int16 / uint16 stored activity only")
            return engine.to_device_f32(m, self.device)
        slope, intercept = (1.0, 0.0) if rescale is None else (float(rescale[0]), float(rescale[1]))
        if isinstance(m, torch.Tensor):
            unsigned = m.dtype == torch.uint16
            raw = m.view(torch.int16).to(self.device, non_blocking=True).contiguous()
        else:
            unsigned = m.dtype == np.uint16
            a = np.ascontiguousarray(m)
            raw = engine.HostStager.get(self.device).upload(a) if engine._stageable(a) else \
                torch.from_numpy(a.view(np.int16)).to(self.device, non_blocking=True)
        out = torch.empty(raw.shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            engine.get_lib().i16_to_f32(raw.data_ptr(), unsigned, slope, intercept, out.data_ptr(), raw.numel(),
                                        torch.cuda.current_stream(self.device).cuda_stream)
        return out

    def _convolve(self, maps: Sequence, weights: Optional[Sequence[float]], voxel_size, tissue_densities=None,
                  out: Optional[np.ndarray] = None, ct_hu=None, rescale=None):
        if len(maps) == 0:
            raise ValueError("No activity maps provided")
        shape = tuple(maps[0].shape)
        if len(shape) != 3:
            raise ValueError("activity maps must be 3-D")
        if any(tuple(m.shape) != shape for m in maps):
            raise ValueError("All activity maps must have the same dimensions")
        kdev, tag = self._kernel_for(voxel_size)
        plan = self._plans.get(shape, tuple(kdev.shape), self.boundary, self.device, tag, lambda: kdev, self.algo)
        self._last_plan = plan
        acts = [self._activity_to_device(m, rescale) for m in maps]
        den = None
        if tissue_densities is not None and ct_hu is not None:
            raise ValueError("give tissue_densities or ct_hu, not both")
        if tissue_densities is not None:
            den = engine.to_device_f32(tissue_densities, self.device)
            if tuple(den.shape) != plan.out_shape:
                raise ValueError("tissue_densities must have the shape of the activity map")
        elif ct_hu is not None:
            den = self._density_from_ct(ct_hu, plan.out_shape)
        cfg = self.config
        dose = plan.execute(acts, weights, den, float(cfg.get("rho_ref", 1.0)), float(cfg.get("rho_min", 0.1)),
                            float(cfg.get("rho_cut", 0.0)), float(cfg.get("scale", 1.0)))
        return dose

    HOST_PIPELINE_CHUNKS = 8

    def _host_call(self, maps: Sequence, weights: Optional[Sequence[float]], voxel_size, tissue_densities=None, out=None, ct_hu=None,
                   rescale=None):
        """Host arrays in -> host array out (the reference's calling convention, core/kernel_convolution.py:48-76).
        CUDA tensors in -> CUDA tensor out, no copies.

        The host form is a pipeline over the link, which is what an end-to-end call costs (the convolution itself is
        ~1 ms of a ~20 ms call): the activity goes up through the staging engine and the forward passes and the x / y
        inverse passes run; then, plane block by plane block, the density (or int16 CT) block goes up on one stream, the
        output pass of that block runs, and the finished dose block goes down on a third stream straight into pinned
        memory - upload of block i+1 and download of block i use the two directions of the link at the same time."""
        if len(maps) == 0:
            raise ValueError("No activity maps provided")
        on_device = isinstance(maps[0], torch.Tensor) and maps[0].is_cuda
        den_src = tissue_densities if tissue_densities is not None else ct_hu
        host_den = den_src is None or isinstance(den_src, np.ndarray)
        if on_device or not host_den:
            dose = self._convolve(maps, weights, voxel_size, tissue_densities, ct_hu=ct_hu, rescale=rescale)
            return dose if on_device else self._to_host(dose, out)
        shape = tuple(maps[0].shape)
        if len(shape) != 3:
            raise ValueError("activity maps must be 3-D")
        if any(tuple(m.shape) != shape for m in maps):
            raise ValueError("All activity maps must have the same dimensions")
        if tissue_densities is not None and ct_hu is not None:
            raise ValueError("give tissue_densities or ct_hu, not both")
        kdev, tag = self._kernel_for(voxel_size)
        plan = self._plans.get(shape, tuple(kdev.shape), self.boundary, self.device, tag, lambda: kdev, self.algo)
        self._last_plan = plan
        want64 = str(self.config.get("output_dtype", "float32")) == "float64"
        pinned_result = out is None and not want64
        if plan.info.algo != engine._capi.ALGO_FFT or len(maps) > engine.MAX_T or int(np.prod(shape)) * 4 < (8 << 20):
            dose = self._convolve(maps, weights, voxel_size, tissue_densities, ct_hu=ct_hu, rescale=rescale)
            return self._to_host(dose, out)
        if den_src is not None and tuple(den_src.shape) != plan.out_shape:
            raise ValueError(("tissue_densities" if tissue_densities is not None else "ct_hu") + " must have the shape of the activity map")
        cfg, dev, lib, h = self.config, self.device, plan.lib, plan.handle
        rho_ref, rho_min, rho_cut, scale = (float(cfg.get("rho_ref", 1.0)), float(cfg.get("rho_min", 0.1)), float(cfg.get("rho_cut", 0.0)),
                                            float(cfg.get("scale", 1.0)))
        gain = scale * (rho_ref if den_src is not None else 1.0)
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            acts = [self._activity_to_device(m, rescale) for m in maps]
            w = None if weights is None else [float(x) for x in weights]
            lib.conv_forward_planes(h, [a.data_ptr() for a in acts], w, gain, 0, shape[0], main.cuda_stream)
            lib.conv_middle(h, main.cuda_stream)
            O0 = plan.out_shape[0]
            d_out = torch.empty(plan.out_shape, dtype=torch.float32, device=dev)
            d_den = None
            if den_src is not None:
                den_arr = np.ascontiguousarray(den_src)
                is_hu = ct_hu is not None
                if den_arr.dtype not in (np.float32, np.float64, np.int16):
                    den_arr = den_arr.astype(np.float32)
                d_den = torch.empty(plan.out_shape, dtype=torch.float32, device=dev)
                d_hu = torch.empty(plan.out_shape, dtype=torch.int16, device=dev) if (is_hu and den_arr.dtype == np.int16) else None
                if is_hu:
                    from ..tissue.density import HU_KNOTS

                    knots = cfg.get("hu_knots", HU_KNOTS)
            host = torch.empty(plan.out_shape, dtype=torch.float32, pin_memory=True) if pinned_result else None
            if not hasattr(self, "_side"):
                self._side = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            s_in, s_out = self._side
            s_in.wait_stream(main)
            s_out.wait_stream(main)
            stager = engine.HostStager.get(dev)
            nch = min(self.HOST_PIPELINE_CHUNKS, O0)
            bounds = [O0 * i // nch for i in range(nch + 1)]
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                if d_den is not None:
                    with torch.cuda.stream(s_in):
                        if d_hu is not None:
                            stager.upload(den_arr[lo:hi], out=d_hu[lo:hi])
                            lib.hu_to_density(d_hu[lo:hi].data_ptr(), True, knots, d_den[lo:hi].data_ptr(), d_den[lo:hi].numel(), s_in.cuda_stream)
                        else:
                            stager.upload(den_arr[lo:hi], out=d_den[lo:hi])
                            if is_hu:
                                lib.hu_to_density(d_den[lo:hi].data_ptr(), False, knots, d_den[lo:hi].data_ptr(), d_den[lo:hi].numel(), s_in.cuda_stream)
                    main.wait_stream(s_in)
                lib.conv_output_planes(h, None if d_den is None else d_den.data_ptr(), rho_min, rho_cut, d_out.data_ptr(), lo, hi, main.cuda_stream)
                if host is not None:
                    s_out.wait_stream(main)
                    with torch.cuda.stream(s_out):
                        host[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
            if host is not None:
                s_out.synchronize()
                main.wait_stream(s_out)
                res = host.numpy()
            else:
                res = engine.to_host(d_out, out, want64)
            main.synchronize()  # every tensor of this call is idle before the allocator may hand it out again
        self._check_device_errors()
        return out if (out is not None and not isinstance(out, torch.Tensor)) else res

    def _to_host(self, dose: torch.Tensor, out=None) -> np.ndarray:
        want64 = str(self.config.get("output_dtype", "float32")) == "float64"
        host = engine.to_host(dose, out, want64)
        self._check_device_errors()
        return out if (out is not None and not isinstance(out, torch.Tensor)) else host

    def _check_device_errors(self) -> None:
        """Every host-returning path ends here: read (and clear) the device-side TMA watchdog flags of the plans this
        calculator has used and raise PvdoseError rather than hand back a dose map computed from a tile that never
        arrived.  Device-tensor-returning calls stay asynchronous: call this after your own synchronisation."""
        plan = getattr(self, "_last_plan", None)
        if plan is not None and plan.handle:
            plan.check_device_errors()

    # ------------------------------------------------------------------ reference API
    def calculate_dose_rate(self, activity_map, voxel_size: Tuple[float, float, float] = None,
                            tissue_densities=None, out=None, ct_hu=None, rescale=None):
        """A1.  Host ndarray in -> host ndarray out; CUDA tensor in -> CUDA tensor out (no copies).
        Density correction (A9): `tissue_densities` (g/cm3) or `ct_hu` (the CT in Hounsfield units, int16 or float).
        `activity_map` may be float64 / float32, or int16 / uint16 stored values with `rescale` = (slope, intercept)."""
        return self._host_call([activity_map], None, voxel_size, tissue_densities, out, ct_hu, rescale)

    def calculate_dose_rate_batch(self, activity_maps: Sequence, voxel_size=None, tissue_densities=None,
                                  outs: Optional[Sequence] = None, ct_hu=None, rescale=None) -> list:
        """A1 for a batch of independent host volumes (one per patient / time point), software pipelined:
        the H2D copy of volume i+1, the convolution of volume i and the D2H copy of volume i-1 run on three
        CUDA streams with double-buffered device tensors, so the PCIe link (the end-to-end bottleneck: the
        convolution itself is ~1 ms) works in both directions at once.  `tissue_densities` is None, one
        volume shared by all, or one per activity map.  `outs`: optional host tensors/arrays to fill (pinned
        host tensors avoid a staging copy; they may repeat, e.g. two alternating buffers).  `ct_hu`: instead of
        densities, one int16 CT volume (Hounsfield units) per activity map (or one shared); it is copied as int16
        (2 bytes per voxel over the link) and turned into density on the device.  int16 / uint16 activity volumes
        (stored PET values) are copied as 16-bit and become slope * stored + intercept on the device (`rescale`)."""
        n = len(activity_maps)
        if ct_hu is not None:
            if tissue_densities is not None:
                raise ValueError("give tissue_densities or ct_hu, not both")
            from ..tissue.density import HU_KNOTS

            knots = self.config.get("hu_knots", HU_KNOTS)
            if not isinstance(ct_hu, (list, tuple)):
                tissue_densities, ct_hu = self._density_from_ct(ct_hu, tuple(activity_maps[0].shape)) if n else None, None
        if n == 0:
            return []
        shape = tuple(activity_maps[0].shape)
        if any(tuple(a.shape) != shape for a in activity_maps):
            raise ValueError("All activity maps must have the same dimensions")
        per_vol_ct = ct_hu is not None
        if per_vol_ct:
            if len(ct_hu) != n:
                raise ValueError("need one CT volume per activity map")
            tissue_densities = ct_hu
        per_vol_den = isinstance(tissue_densities, (list, tuple))
        if per_vol_den and len(tissue_densities) != n:
            raise ValueError("need one density volume per activity map")
        kdev, tag = self._kernel_for(voxel_size)
        plan = self._plans.get(shape, tuple(kdev.shape), self.boundary, self.device, tag, lambda: kdev, self.algo)
        cfg, dev = self.config, self.device
        rr, rm, rc, sc = (float(cfg.get("rho_ref", 1.0)), float(cfg.get("rho_min", 0.1)), float(cfg.get("rho_cut", 0.0)),
                          float(cfg.get("scale", 1.0)))

        def host_f32(x):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            return t if t.dtype == torch.float32 else t.to(torch.float32)

        def host_i16(x):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            return t if t.dtype == torch.int16 else t.to(torch.int16)

        s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream(dev))
        for s in (s_in, s_cmp, s_out):
            s.wait_event(start)
        d_act = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(2)]
        first = activity_maps[0]
        act16 = (first.dtype in (torch.int16, torch.uint16)) if isinstance(first, torch.Tensor) else (np.asarray(first).dtype in (np.int16, np.uint16))
        act_unsigned = act16 and ((first.dtype == torch.uint16) if isinstance(first, torch.Tensor) else (np.asarray(first).dtype == np.uint16))
        slope, intercept = (1.0, 0.0) if rescale is None else (float(rescale[0]), float(rescale[1]))
        if rescale is not None and not act16:
            raise ValueError("rescale=(slope, intercept) applies to int16 / uint16 stored activity only")
        d_raw = [torch.empty(shape, dtype=torch.int16, device=dev) for _ in range(2)] if act16 else None

        def host_raw16(x):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x).view(np.int16))
            return t.view(torch.int16)
        d_out = [torch.empty(plan.out_shape, dtype=torch.float32, device=dev) for _ in range(2)]
        d_den = None
        if tissue_densities is not None:
            if per_vol_den:
                d_den = [torch.empty(plan.out_shape, dtype=torch.float32, device=dev) for _ in range(2)]
                if per_vol_ct:
                    d_hu = [torch.empty(plan.out_shape, dtype=torch.int16, device=dev) for _ in range(2)]
            else:
                shared = engine.to_device_f32(tissue_densities, dev)
                start2 = torch.cuda.Event()
                start2.record(torch.cuda.current_stream(dev))
                s_cmp.wait_event(start2)
        ev_in = [torch.cuda.Event() for _ in range(n)]
        ev_cmp = [torch.cuda.Event() for _ in range(n)]
        ev_out = [torch.cuda.Event() for _ in range(n)]
        results = []
        for i in range(n):
            slot = i & 1
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_cmp[i - 2])  # device input slot free again
                if act16:
                    d_raw[slot].copy_(host_raw16(activity_maps[i]), non_blocking=True)
                else:
                    d_act[slot].copy_(host_f32(activity_maps[i]), non_blocking=True)
                if d_den is not None and per_vol_ct:
                    h = host_i16(tissue_densities[i])
                    if tuple(h.shape) != plan.out_shape:
                        raise ValueError("ct_hu must have the shape of the activity map")
                    d_hu[slot].copy_(h, non_blocking=True)
                elif d_den is not None:
                    d_den[slot].copy_(host_f32(tissue_densities[i]), non_blocking=True)
                ev_in[i].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[i])
                if i >= 2:
                    s_cmp.wait_event(ev_out[i - 2])  # device output slot drained
                den = None if tissue_densities is None else (d_den[slot] if d_den is not None else shared)
                if act16:  # stored 16-bit activity -> float32 on the device (d_act[slot] is free: ev_cmp[i - 2] has been waited for)
                    with torch.cuda.device(dev):
                        engine.get_lib().i16_to_f32(d_raw[slot].data_ptr(), act_unsigned, slope, intercept, d_act[slot].data_ptr(),
                                                    d_act[slot].numel(), torch.cuda.current_stream(dev).cuda_stream)
                if per_vol_ct:  # HU -> density on the device, in the compute stream (d_den[slot] is free: see ev_out wait)
                    self._lib_hu_to_density(d_hu[slot], knots, d_den[slot])
                plan.execute([d_act[slot]], None, den, rr, rm, rc, sc, out=d_out[slot])
                ev_cmp[i].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[i])
                if outs is not None:
                    tgt = outs[i] if isinstance(outs[i], torch.Tensor) else torch.from_numpy(outs[i])
                else:
                    tgt = torch.empty(plan.out_shape, dtype=torch.float32, pin_memory=True)
                tgt.copy_(d_out[slot], non_blocking=True)
                ev_out[i].record(s_out)
                results.append(tgt)
        for s in (s_in, s_cmp, s_out):
            s.synchronize()
        plan.check_device_errors()
        return [r.numpy() for r in results]

    def _lib_hu_to_density(self, hu: torch.Tensor, knots, rho: torch.Tensor) -> None:
        with torch.cuda.device(self.device):
            engine.get_lib().hu_to_density(hu.data_ptr(), hu.dtype == torch.int16, knots, rho.data_ptr(), hu.numel(),
                                           torch.cuda.current_stream(self.device).cuda_stream)

    def calculate_absorbed_dose(self, activity_maps, time_points: List[float], voxel_size=None,
                                tissue_densities=None, out=None, ct_hu=None):
        """A2.  time_points in hours; dose = sum_i w_i * conv(a_i, k), w = trapezoid weights * 3600."""
        if len(activity_maps) != len(time_points):
            raise ValueError("Number of activity maps must match number of time points")
        w = trapezoid_weights(time_points, HOURS_TO_SECONDS)
        return self._host_call(list(activity_maps), w, voxel_size, tissue_densities, out, ct_hu)

    def calculate_absorbed_dose_from_accumulated(self, accumulated_activity, voxel_size=None, tissue_densities=None, out=None):
        """Called by the reference front door (core/dose_calculator.py:104,123) but defined nowhere there;
        meaning: A1 applied to time-integrated activity (Bq*time -> dose)."""
        return self.calculate_dose_rate(accumulated_activity, voxel_size, tissue_densities, out)

    def calculate_weighted(self, activity_maps, weights: Sequence[float], voxel_size=None, tissue_densities=None, out=None):
        """conv(sum_i w_i a_i, k) for caller-chosen weights (used by DoseCalculator's activity mode)."""
        return self._host_call(list(activity_maps), [float(x) for x in weights], voxel_size, tissue_densities, out)

    def _resample_activity(self, activity_map, input_voxel_size, output_voxel_size):
        """The reference stub returns None (kernel_convolution.py:108-115).  Not needed here: the kernel is
        evaluated on the image grid instead (A10)."""
        raise NotImplementedError("activity resampling is replaced by evaluating the kernel on the image grid")
