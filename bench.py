#!/usr/bin/env python
"""Benchmark of the kernel-convolution dose path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1|c4|c5] [--boundary reference|same]

A *step* is one pass of the hot path over one synthetic patient volume per GPU:
  workload c3 (default, the configuration the metric is quoted on): 512x512x400 float32 activity,
  Y90/water 51^3 dose voxel kernel @ 1 mm, voxel-wise density correction, reference boundary mode.
Rank 0 prints ONE JSON line.  `value` = whole-job dose volumes/s with inputs resident in HBM;
`e2e` = the same through the public Python calculator API with pinned HOST buffers (H2D + D2H inside
the timed region).  N > 1 (torchrun, one rank per GPU, NCCL): independent volumes are sharded over the
ranks with no data-path collective (weak scaling).  `--impl reference` times the reference's own CPU
path (the literal np.fft expression of core/kernel_convolution.py:71-74, single-threaded by
construction) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
# stdout carries exactly ONE JSON line.  Native libraries write there too (NCCL prints its version banner on stdout
# when NCCL_DEBUG is VERSION/INFO in the environment or in a nccl.conf), so file descriptor 1 is pointed at stderr for
# the whole run and the JSON line goes to the saved, real stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


WORKLOADS = {
    # name: (shape, kernel grid, nuclide, voxel mm, T, density)
    "c3": dict(shape=(512, 512, 400), kgrid=(51, 51, 51), nuclide="Y90", voxel=1.0, T=1, density=True,
               desc="C3: Y90 whole-body 512x512x400, 51^3 DVK @1mm, density-corrected"),
    "c2": dict(shape=(256, 256, 256), kgrid=(31, 31, 31), nuclide="Lu177", voxel=4.8, T=4, density=False,
               desc="C2: Lu-177 4-timepoint time-integrated dose 256^3, 31^3 DVK @4.8mm"),
    "c1": dict(shape=(48, 48, 48), kgrid=(64, 64, 64), nuclide="Y90", voxel=1.0, T=1, density=False,
               desc="C1: examples/single_timepoint_y90_physical_decay.py 48^3, 64^3 DVK"),
    # multi-GPU configurations (strong scaling: the job is fixed, ranks share it)
    "c4": dict(shape=(256, 256, 256), kgrid=(31, 31, 31), nuclide="Lu177", voxel=4.8, T=4, density=False, volumes=64,
               desc="C4: 64 patient volumes (256^3, 4 time points, 31^3 DVK) sharded over the ranks"),
    "c5": dict(shape=(1024, 1024, 800), kgrid=(51, 51, 51), nuclide="Y90", voxel=1.0, T=1, density=True,
               desc="C5: 1024x1024x800 volume, 51^3 DVK, slabs along axis 0 + kernel-radius halo exchange over NVLink"),
}


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def synth_inputs(wl, seed=90):
    """Synthetic activity / density of SURVEY.md section 8d (fixed seed), float32."""
    rng = np.random.default_rng(seed)
    n0, n1, n2 = wl["shape"]
    acts = []
    base = rng.uniform(0.0, 1e2, size=wl["shape"]).astype(np.float32)
    sl = tuple(slice(int(n * 0.39), int(n * 0.39) + max(1, int(n * 0.2))) for n in wl["shape"])
    base[sl] = 2e6
    times = [4.0, 24.0, 96.0, 168.0][: wl["T"]] if wl["T"] > 1 else [2.0]
    for t in times:
        acts.append(base if wl["T"] == 1 else (base * np.float32(np.exp(-np.log(2) * t / 161.52))).astype(np.float32))
    rho = None
    if wl["density"]:
        x = (np.arange(n0, dtype=np.float32) - n0 / 2) / (0.42 * n0)
        y = (np.arange(n1, dtype=np.float32) - n1 / 2) / (0.30 * n1)
        body = (x[:, None] ** 2 + y[None, :] ** 2) <= 1.0
        lung = (((x[:, None] - 0.45) / 0.3) ** 2 + (y[None, :] / 0.5) ** 2 <= 1.0) | (((x[:, None] + 0.45) / 0.3) ** 2 + (y[None, :] / 0.5) ** 2 <= 1.0)
        spine = (x[:, None] ** 2 + ((y[None, :] - 0.6) ** 2)) <= 0.02
        plane = np.full((n0, n1), 0.00129, dtype=np.float32)
        plane[body] = 1.04
        plane[lung & body] = 0.26
        plane[spine] = 1.42
        rho = np.ascontiguousarray(np.broadcast_to(plane[:, :, None], wl["shape"])).astype(np.float32)
    return acts, times, rho


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        hot = sorted(sm)[len(sm) // 2:] if sm else []  # upper half ~ samples under load
        return {"sm_mhz": statistics.median(hot) if hot else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa(local: int, nlocal: int = 1) -> str:
    """One rank per GPU: run this rank's host threads (and so its pinned staging buffers, first touch) on the CPUs
    of the GPU's own NUMA node, so that host<->device copies of different ranks do not share one socket's memory
    controllers / inter-socket link.  Ranks whose GPUs hang off the same node split that node's CPUs between them, so
    that the staging threads of the end-to-end leg (8 per rank at most) do not oversubscribe the cores.  Best effort;
    returns a note for stderr."""
    try:
        import torch

        def cpulist(idx):
            p = torch.cuda.get_device_properties(idx)
            bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
            with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
                spec = f.read().strip()
            cpus = set()
            for part in spec.split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            return bdf, spec, cpus

        bdf, spec, cpus = cpulist(local)
        cpus &= os.sched_getaffinity(0)
        if cpus:
            sharers = [r for r in range(nlocal) if cpulist(r)[2] & cpus] if nlocal > 1 else [local]
            mine = sorted(cpus)
            if len(sharers) > 1 and len(mine) >= 2 * len(sharers):
                k = sharers.index(local)
                per = len(mine) // len(sharers)
                mine = mine[k * per : (k + 1) * per]
            os.sched_setaffinity(0, set(mine))
            return f"rank on GPU {local} ({bdf}) bound to {len(mine)} of the {len(cpus)} local CPUs ({spec}; {len(sharers)} rank(s) share the node)"
        return f"GPU {local} ({bdf}): no usable local CPUs in {spec!r}"
    except Exception as e:  # sysfs layout / permissions differ: keep the default affinity
        return f"NUMA binding skipped: {e}"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
def cpu_threaded_time(wl, acts, rho, kernel64, pick):
    """the same mathematics with every host thread - scipy.fft real transforms,
    workers = all cores (oracle.conv_reference_fast) - on the same sub-volume as the literal run.  Best of 2.
    Returns (volumes/s, cores)."""
    from oracle import dose_oracle as orc

    cores = len(os.sched_getaffinity(0))
    subs = [np.ascontiguousarray(m[: pick[0], : pick[1], : pick[2]]).astype(np.float64) for m in acts]
    r = None if rho is None else rho[: pick[0], : pick[1], : pick[2]]
    tw = [4.0, 24.0, 96.0, 168.0][: len(subs)]
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        if len(subs) == 1:
            d = orc.conv_reference_fast(subs[0], kernel64, workers=cores)
        else:
            d = orc.absorbed_dose_trapezoid(subs, tw, kernel64, conv=lambda a, k: orc.conv_reference_fast(a, k, workers=cores))
        if r is not None:
            d = orc.density_correct(d, r)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return subs[0].size / best / float(np.prod(wl["shape"])), cores


def reference_step(maps64, kernel64, rho):
    """One unit of the reference's work for the workload: T = 1 -> calculate_dose_rate, T > 1 -> the literal
    calculate_absorbed_dose loop (one FFT convolution per time point, kernel FFT recomputed every time)."""
    from oracle import dose_oracle as orc

    if len(maps64) == 1:
        d = orc.conv_reference(maps64[0], kernel64)
    else:
        d = orc.absorbed_dose_trapezoid(maps64, [4.0, 24.0, 96.0, 168.0][: len(maps64)], kernel64)
    if rho is not None:
        d = orc.density_correct(d, rho)
    return d


def load_reference_calculator(wl):
    """The reference's OWN class (core/kernel_convolution.py:26-76) from oracle/_ref (a byte-identical copy of
    /root/reference made by oracle/ref_loader.py; it travels to the GPU box), with the workload's dose voxel kernel
    produced by the reference's own generator and assigned through the public mutable attribute `calc.kernel`
    (the class hard-codes a 64^3 grid, kernel_convolution.py:45).  The one NaN the Y90 / Ga68 generators leave at r = 0
    (y90_kernel.py:134-138) is replaced by the finite centre value, as in SURVEY.md Appendix C.4 - with it every dose
    voxel of the reference is NaN.  Returns (calculator, float64 kernel) or raises when oracle/_ref is absent."""
    from oracle import dose_oracle as orc
    from oracle import ref_loader

    ref_loader.import_reference()
    from pyvoxeldosimetry.core.kernel_convolution import KernelConvolutionCalculator as RefCalc
    from pyvoxeldosimetry.data.dose_kernels.kernel_factory import KernelFactory as RefFactory

    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        k = RefFactory().get_kernel(wl["nuclide"], "water", voxel_size=wl["voxel"], grid_size=tuple(wl["kgrid"]), force_regenerate=True)
        calc = RefCalc(wl["nuclide"], "water", kernel_resolution=wl["voxel"])
    k = np.array(k, dtype=np.float64)
    bad = ~np.isfinite(k)
    if bad.any():
        k[bad] = orc.make_kernel(wl["nuclide"], wl["voxel"], wl["kgrid"])[bad]
    calc.kernel = k
    return calc, k


def reference_step_real(calc, maps64, times, vox, rho):
    """One unit of the reference's work through its own public API: T = 1 -> calculate_dose_rate, T > 1 ->
    calculate_absorbed_dose (core/kernel_convolution.py:48-106); the density correction (absent from the reference,
    core/dose_calculator.py:90) is the float64 NumPy formula of SURVEY section 8 A9 applied to its result."""
    from oracle import dose_oracle as orc

    if len(maps64) == 1:
        d = calc.calculate_dose_rate(activity_map=maps64[0], voxel_size=vox)
    else:
        d = calc.calculate_absorbed_dose(activity_maps=maps64, time_points=times, voxel_size=vox)
    if rho is not None:
        d = orc.density_correct(d, rho)
    return d


def arm_config(wl, boundary):
    """The `config` object of the bench line - IDENTICAL for both arms (`--impl ours` / `--impl reference`) of one workload, so
    that the driver can see that they measured the same thing; what differs between the arms goes into `impl_config`."""
    nbytes = 4 * int(np.prod(wl["shape"])) * (wl["T"] + (1 if wl["density"] else 0))
    return {"workload": wl["desc"], "boundary": boundary, "volumes_per_step_per_gpu": 1,
            "l2": (f"inputs ({nbytes / 1e6:.0f} MB per step) exceed the 126 MB L2, nothing to flush" if nbytes > 126e6 else
                   "small workload: inputs fit the 126 MB L2 (launch-bound), not flushed")}


def run_reference(args, wl):
    """`--impl reference`: the UNMODIFIED reference class on the full workload shape, on this box's host cores
    (np.fft is single-threaded by construction: cores = 1).  Every step is one full volume: K steps + W warm-ups of
    ~10 s each for C3.  Falls back to the oracle port (kind "port") only when oracle/_ref was not built."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    from oracle import dose_oracle as orc

    acts, times, rho = synth_inputs(wl)
    nvox = float(np.prod(wl["shape"]))
    maps64 = [a.astype(np.float64) for a in acts]
    vox = (float(wl["voxel"]),) * 3
    del acts
    try:
        calc, k64 = load_reference_calculator(wl)
        kind = "reference"
        what = ("pyvoxeldosimetry.core.kernel_convolution.KernelConvolutionCalculator." +
                ("calculate_dose_rate" if wl["T"] == 1 else "calculate_absorbed_dose") + " of oracle/_ref (byte-identical copy of the reference)")
        step = lambda: reference_step_real(calc, maps64, times, vox, rho)
    except Exception as e:  # oracle/_ref not built (no /root/reference at build time): literal port of the same four lines
        sys.stderr.write(f"[bench] reference package unavailable ({e}); timing the oracle port\n")
        k64 = orc.make_kernel(wl["nuclide"], wl["voxel"], wl["kgrid"]).astype(np.float32).astype(np.float64)
        kind, what = "port", "oracle.conv_reference (literal np.fft expression of core/kernel_convolution.py:71-74)"
        step = lambda: reference_step(maps64, k64, rho)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d = step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = 1.0 / dt
    sample = (f"{what}, float64, full {wl['shape'][0]}x{wl['shape'][1]}x{wl['shape'][2]} volume per step, {args.steps} steps after "
              f"{args.warmup} warm-ups; np.fft is single-threaded" + ("; + float64 NumPy density correction" if rho is not None else ""))
    line = {
        "impl": "reference", "metric": "dose_volumes_per_sec", "value": value, "unit": "volumes/s", "voxels_per_sec": value * nvox,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": arm_config(wl, args.boundary),
        "impl_config": {"parallelism": "one host process, one thread (the reference has no parallelism)"},
        "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(), "finite": bool(np.isfinite(d).all())},
        "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def link_peak(world: int):
    """Bare pinned cudaMemcpyAsync ceiling of the box measured by scripts/link_peak.py (profiles/r02_link_peak_<N>gpu.json):
    per-rank H2D / D2H GB/s when all `world` ranks copy at once.  None when no file exists for this N."""
    p = os.path.join(REPO, "profiles", f"r02_link_peak_{world}gpu.json")
    try:  # an evidence file must never cost the bench line: tolerate banners before the JSON object, or a bad file
        lines = [ln for ln in open(p).read().splitlines() if ln.startswith("{")]
        d = json.loads(lines[-1])
    except Exception:
        return None
    return {"h2d_gbs_per_rank": d["h2d_GBs_aggregate"] / world, "d2h_gbs_per_rank": d["d2h_GBs_aggregate"] / world,
            "concurrent_gbs_per_rank_each_way": d["h2d_d2h_concurrent_GBs_aggregate_each_way"] / world, "source": os.path.basename(p)}


def measure_e2e(wl, args, calc, plan, acts_h, rho_h, dev, dist, world, barrier):
    """The same metric end to end: HOST buffers in, HOST dose map out, every copy inside the timed region.
    Headline = the reference's own calling convention: DoseCalculator.calculate_dose with pageable
    NumPy arrays, one blocking call per volume, the result a fresh ndarray.  The other variants (16-bit inputs as
    scanners store them, pinned buffers, the pipelined batch call) are reported beside it in variants_ms."""
    import torch

    from pyvoxeldosimetry_b200 import DoseCalculator

    vox = (float(wl["voxel"]),) * 3
    shape = wl["shape"]
    nvox = int(np.prod(shape))
    reps = max(3, min(args.steps, 8))
    front = DoseCalculator(wl["nuclide"], "kernel", {"kernel_grid": wl["kgrid"], "boundary": args.boundary, "device": str(dev),
                                                      "kernel_resolution": wl["voxel"], "tissue_name": "water"})
    a32 = acts_h[0]
    variants, bytes_of = {}, {}

    def timed(name, fn, h2d, d2h, n_per_call=1):
        for _ in range(3):  # warm-up holds its result like the timed loop does: both pinned result blocks get cached
            r = fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / (reps * n_per_call)
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        variants[name] = float(t.item()) * 1e3
        bytes_of[name] = (int(h2d), int(d2h))
        return r

    if wl["T"] == 1:
        den = {"tissue_densities": rho_h} if rho_h is not None else {}
        nden = nvox * 4 if rho_h is not None else 0
        head = "front_door_f32_ndarray" + ("_density_f32_ndarray" if rho_h is not None else "")
        res = timed(head, lambda: front.calculate_dose(activity_maps=[a32], time_points=[2.0], voxel_size=vox, **den).dose_rate_maps[0],
                    nvox * 4 + nden, nvox * 4)
        assert isinstance(res, np.ndarray) and res.shape == tuple(shape)
        a64 = a32.astype(np.float64)
        timed("front_door_f64_ndarray" + ("_density_f32_ndarray" if rho_h is not None else ""),
              lambda: front.calculate_dose(activity_maps=[a64], time_points=[2.0], voxel_size=vox, **den).dose_rate_maps[0], nvox * 4 + nden, nvox * 4)
        del a64
        if rho_h is not None:
            timed("front_door_f32_ndarray_no_density", lambda: front.calculate_dose(activity_maps=[a32], time_points=[2.0], voxel_size=vox).dose_rate_maps[0],
                  nvox * 4, nvox * 4)
            # the CT as scanners store it: int16 Hounsfield units (2 bytes per voxel over the link), density derived on the device
            # (HU -1000 / -700 / 32 / 350 <-> 0.00129 / 0.26 / 1.04 / 1.42 g/cm3 through tissue.HU_KNOTS)
            hu_h = np.full(shape, -1000, dtype=np.int16)
            hu_h[rho_h > 0.2] = -700
            hu_h[rho_h > 1.0] = 32
            hu_h[rho_h > 1.4] = 350
            kc = front.calculator
            timed("calculator_f32_ndarray_ct_i16_ndarray", lambda: kc.calculate_dose_rate(a32, vox, ct_hu=hu_h), nvox * 6, nvox * 4)
            # PET as DICOM stores it: 16-bit values + rescale slope (io/dicom.py:27-47)
            slope = float(a32.max()) / 32000.0
            st16 = np.clip(np.rint(a32 / slope), 0, 32767).astype(np.int16)
            timed("calculator_i16_ndarray_rescale_ct_i16_ndarray", lambda: kc.calculate_dose_rate(st16, vox, ct_hu=hu_h, rescale=(slope, 0.0)), nvox * 4, nvox * 4)
            # pinned buffers + the pipelined batch call (H2D of volume i+1, convolution of i, D2H of i-1 overlap)
            nb = 8
            pin = lambda x: torch.from_numpy(x).pin_memory()
            p_a, p_rho, p_hu, p_st = pin(a32), pin(rho_h), pin(hu_h), pin(st16)
            outs = [torch.empty(tuple(shape), dtype=torch.float32).pin_memory() for _ in range(2)]
            bo = [outs[i & 1] for i in range(nb)]
            timed("batch_pinned_f32_density_f32", lambda: kc.calculate_dose_rate_batch([p_a] * nb, vox, [p_rho] * nb, bo), nvox * 8, nvox * 4, nb)
            timed("batch_pinned_f32_ct_i16", lambda: kc.calculate_dose_rate_batch([p_a] * nb, vox, None, bo, ct_hu=[p_hu] * nb), nvox * 6, nvox * 4, nb)
            timed("batch_pinned_i16_rescale_ct_i16", lambda: kc.calculate_dose_rate_batch([p_st] * nb, vox, None, bo, ct_hu=[p_hu] * nb, rescale=(slope, 0.0)),
                  nvox * 4, nvox * 4, nb)
            del p_a, p_rho, p_hu, p_st, outs, bo
        api = ("DoseCalculator.calculate_dose(activity_maps=[float32 ndarray], time_points=[t], voxel_size, tissue_densities=float32 ndarray)"
               ".dose_rate_maps[0] -> ndarray; pageable NumPy arrays, one blocking call per volume")
    else:
        kc = front.calculator
        times = [4.0, 24.0, 96.0, 168.0][: wl["T"]]
        head = f"calculator_absorbed_dose_{wl['T']}x_f32_ndarray"
        timed(head, lambda: kc.calculate_absorbed_dose(acts_h, times, vox), nvox * 4 * wl["T"], nvox * 4)
        api = "KernelConvolutionCalculator.calculate_absorbed_dose(list of float32 ndarrays, time_points, voxel_size) -> ndarray"
    ms = variants[head]
    h2d, d2h = bytes_of[head]
    lp = link_peak(world)
    out = {"value": world / (ms * 1e-3), "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms, "api": api,
           "headline_variant": head, "variants_ms": {k: round(v, 3) for k, v in variants.items()},
           "variants_bytes": {k: {"h2d": b[0], "d2h": b[1]} for k, b in bytes_of.items()}}
    if lp:
        # a blocking call cannot overlap its own upload and download: ideal = H2D bytes / H2D peak + D2H bytes / D2H peak
        ideal = (h2d / lp["h2d_gbs_per_rank"] + d2h / lp["d2h_gbs_per_rank"]) / 1e6
        out["link"] = dict(lp, ideal_ms_serial=round(ideal, 3), link_frac=round(ideal / ms, 3))
        best_b = min((k for k in variants if k.startswith("batch_")), key=lambda k: variants[k], default=None)
        if best_b:
            bh, bd = bytes_of[best_b]
            idealb = max(bh, bd) / lp["concurrent_gbs_per_rank_each_way"] / 1e6  # pipelined: both directions at once
            out["link"]["best_pipelined_variant"] = {"name": best_b, "ms": round(variants[best_b], 3), "ideal_ms": round(idealb, 3),
                                                     "link_frac": round(idealb / variants[best_b], 3)}
    front.calculator._plans.clear()
    return out


def time_steps(step, steps, warmup, dev, dist):
    """CUDA-event time of `steps` calls after `warmup`, barrier + synchronize on both sides, max over ranks (ms per step)."""
    import torch

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_c4(args, dev, dist, rank, world, steps):
    """C4: 64 independent patient volumes (256^3, 4 time points, 31^3 Lu-177 kernel) sharded over the ranks, no collective."""
    import torch

    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator
    from pyvoxeldosimetry_b200.core import trapezoid_weights
    from pyvoxeldosimetry_b200.engine import ConvPlan
    from pyvoxeldosimetry_b200.multi_gpu import shard_range

    wl = WORKLOADS["c4"]
    calc = KernelConvolutionCalculator(wl["nuclide"], "water", wl["voxel"], config={"kernel_grid": wl["kgrid"], "device": str(dev)})
    mine = shard_range(wl["volumes"], world, rank)
    plan = ConvPlan(wl["shape"], wl["kgrid"], "reference", dev)
    plan.set_kernel(calc._kernel_dev)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    sets = [[torch.rand(wl["shape"], device=dev, generator=g) for _ in range(wl["T"])] for _ in range(2)]  # two patients' buffers, alternated
    w = trapezoid_weights([4.0, 24.0, 96.0, 168.0], 3600.0)
    out = torch.empty(plan.out_shape, device=dev)

    batch_acts, batch_outs = [sets[v & 1] for v in mine], [out] * len(mine)

    def step():  # the rank's whole share in ONE batched C-ABI call (pvd_conv_execute_batch)
        plan.execute_batch(batch_acts, w, None, outs=batch_outs)

    ms = time_steps(step, steps, 3, dev, dist)
    plan.check_device_errors()
    plan.close()
    nvox = float(np.prod(wl["shape"]))
    alg = 4.0 * (wl["T"] + 1) * nvox * wl["volumes"]
    peak_gbs, peak_src = peaks()
    return {"workload": wl["desc"], "scaling": "strong", "patients": wl["volumes"], "patients_per_rank": len(mine), "ms_per_job": ms,
            "patients_per_sec": wl["volumes"] * 1e3 / ms, "voxels_per_sec": wl["volumes"] * nvox * 1e3 / ms, "gpu_launches_per_job": 5 * len(mine),
            "roofline": {"bound": "hbm", "achieved": round(alg / (ms * 1e-3) / 1e9, 1), "peak": peak_gbs * world, "unit": "GB/s",
                         "frac": round(alg / (ms * 1e-3) / 1e9 / (peak_gbs * world), 4), "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "what": "algorithmic 4*(T+1) B/voxel * 64 patients / job time"}}


def run_c5(args, dev, dist, rank, world, steps, boundary):
    """C5: one 1024x1024x800 volume, 51^3 Y90 kernel, density-corrected.  N = 1: the whole volume on one GPU (the strong-
    scaling anchor).  N > 1: slabs along axis 0, kernel-radius halo exchange over NVLink on a side stream (peer-memory
    pulls by the copy engines, NCCL send/recv where the ranks cannot map each other), hidden behind the plane-local passes of
    the rank's own planes; the stitched slab result is compared with a single-GPU
    convolution of the whole volume on rank 0 (parity of the NCCL data path, inside the driver-run bench)."""
    import torch

    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator
    from pyvoxeldosimetry_b200.engine import ConvPlan
    from pyvoxeldosimetry_b200.multi_gpu import SlabConvolver

    wl = WORKLOADS["c5"]
    shape = wl["shape"]
    nvox = float(np.prod(shape))
    calc = KernelConvolutionCalculator(wl["nuclide"], "water", wl["voxel"], config={"kernel_grid": wl["kgrid"], "boundary": boundary, "device": str(dev)})
    kdev = calc._kernel_dev
    peak_gbs, peak_src = peaks()
    alg = 12.0 * nvox
    res = {"workload": wl["desc"], "boundary": boundary, "scaling": "strong", "n_gpus": world}
    # the same global volume on every rank (same seed), built plane block by plane block to bound the transient memory
    def global_planes(lo, hi, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        full = torch.rand(shape, device=dev, generator=g)
        return full[lo:hi].clone()

    if world == 1:
        plan = ConvPlan(shape, wl["kgrid"], boundary, dev)
        plan.set_kernel(kdev)
        a = global_planes(0, shape[0], 5)
        rho = global_planes(0, shape[0], 6) + 0.5
        out = torch.empty(plan.out_shape, device=dev)
        ms = time_steps(lambda: plan.execute([a], None, rho, out=out), steps, 3, dev, None)
        plan.check_device_errors()
        res.update({"decomposition": "whole volume on one GPU", "fft_shape": list(plan.fft_shape), "gpu_launches_per_volume": 5})
        plan.close()
    else:
        sc = SlabConvolver(shape, kdev, boundary, device=dev)
        sc.interior.copy_(global_planes(sc.lo, sc.hi, 5))
        rho = global_planes(sc.lo, sc.hi, 6) + 0.5
        ms = time_steps(lambda: sc(density_slab=rho), steps, 3, dev, dist)
        ms_serial = time_steps(lambda: sc(density_slab=rho, overlap=False), max(3, steps // 2), 2, dev, dist)
        ms_compute = time_steps(lambda: sc(density_slab=rho, exchange=False), max(3, steps // 2), 2, dev, dist)
        sc.check_device_errors()
        halo = sc.geom["n"][0] - (sc.hi - sc.lo)
        how = ("halo planes pulled from the neighbours' mapped buffers by copy-engine peer copies (CUDA IPC, stream-ordered flag handshake, no SM)"
               if sc.transport == "peer" else "NCCL send/recv in one group, 32 SMs reserved for it")
        res.update({"decomposition": f"{world} slabs along axis 0 + halo exchange on a side stream under the forward passes of the own planes: {how}",
                    "halo_transport": sc.transport,
                    "slab_planes": sc.hi - sc.lo, "halo_planes": halo, "halo_bytes_received_per_rank": int(halo * shape[1] * shape[2] * 4),
                    "local_fft_shape": list(sc.plan.fft_shape), "ms_exchange_not_overlapped": round(ms_serial, 4),
                    "ms_without_exchange": round(ms_compute, 4), "exchange_exposed_ms": round(ms - ms_compute, 4), "gpu_launches_per_volume": 7})
        # parity of the NCCL data path: stitched slabs vs one whole-volume convolution on rank 0
        mine = sc(density_slab=rho).clone()
        pieces = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        if (shape[0] % world) == 0:
            dist.gather(mine, pieces, dst=0)
        if rank == 0 and pieces is not None:
            del sc
            torch.cuda.empty_cache()
            plan = ConvPlan(shape, wl["kgrid"], boundary, dev)
            plan.set_kernel(kdev)
            whole = plan.execute([global_planes(0, shape[0], 5)], None, global_planes(0, shape[0], 6) + 0.5)
            err = float((torch.cat(pieces) - whole).abs().max() / whole.abs().max())
            plan.close()
            res["slab_vs_whole_volume_max_err_of_peak"] = err
            res["parity_ok"] = bool(err <= 1e-5)
    res.update({"ms_per_volume": ms, "volumes_per_sec": 1e3 / ms, "voxels_per_sec": nvox * 1e3 / ms,
                "roofline": {"bound": "hbm", "achieved": round(alg / (ms * 1e-3) / 1e9, 1), "peak": peak_gbs * world, "unit": "GB/s",
                             "frac": round(alg / (ms * 1e-3) / 1e9 / (peak_gbs * world), 4), "traffic": None,
                             "peak_source": peak_src + f" x {world} GPUs", "what": "algorithmic 12 B/voxel / time per volume"}})
    return res


def dram_traffic_for(build_id: str, workload: str, boundary: str):
    """dram__bytes_read + write of the whole path from the committed ncu capture of this workload - only when the capture was
    taken from THIS build of the kernels (the file carries the library's build id).  c3 / reference: ncu --set full."""
    name = "r02_dram_traffic_c3.json" if (workload, boundary) == ("c3", "reference") else f"r02_dram_traffic_{workload}_{boundary}.json"
    p = os.path.join(REPO, "profiles", name)
    if not os.path.exists(p):
        return None, "no capture for this workload"
    d = json.load(open(p))
    if d.get("build_id") != build_id:
        return None, f"stale: capture is of build {d.get('build_id')}, library is {build_id}"
    return d.get("_whole_path_dram_bytes"), os.path.basename(p)


def run_ours(args, wl):
    import torch

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the dose path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist = None
    full_affinity = os.sched_getaffinity(0)
    # host threads (and so the pinned staging buffers of the e2e leg, first touch) on the GPU's own NUMA node
    sys.stderr.write(bind_to_gpu_numa(local, int(os.environ.get("LOCAL_WORLD_SIZE", world))) + "\n")
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=dev)
    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator
    from pyvoxeldosimetry_b200.engine import ConvPlan

    peak_gbs, peak_src = peaks()
    acts_h, times, rho_h = synth_inputs(wl, seed=90 + rank)
    calc = KernelConvolutionCalculator(wl["nuclide"], "water", wl["voxel"],
                                       config={"kernel_grid": wl["kgrid"], "boundary": args.boundary, "device": str(dev)})
    kdev = calc._kernel_dev
    plan = ConvPlan(wl["shape"], wl["kgrid"], args.boundary, dev)
    plan.set_kernel(kdev)
    acts = [torch.from_numpy(a).to(dev) for a in acts_h]
    rho = None if rho_h is None else torch.from_numpy(rho_h).to(dev)
    out = torch.empty(plan.out_shape, dtype=torch.float32, device=dev)
    w = None
    if wl["T"] > 1:
        from pyvoxeldosimetry_b200.core import trapezoid_weights

        w = trapezoid_weights(times, 3600.0)
    nvox = float(np.prod(wl["shape"]))

    def step():
        plan.execute(acts, w, rho, out=out)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)  # runs from the warm-up to the end of the timed region (nvidia-smi needs ~100 ms per sample)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    # per-kernel device times (CUDA events recorded by the library around each launch), separate pass
    plan.lib.plan_set_profiling(plan.handle, True)
    acc = None
    prof_iters = 5
    for _ in range(prof_iters):
        step()
        pt = plan.lib.plan_get_pass_times(plan.handle)
        if acc is None:
            acc = [[n, 0.0, b] for (n, _, b) in pt]
        for i, (_, ms, _) in enumerate(pt):
            acc[i][1] += ms / prof_iters
    plan.lib.plan_set_profiling(plan.handle, False)
    # ---- timed region: K steps, device resident
    barrier()
    if rank == 0:  # keep the GPU busy long enough for at least a few clock samples under load before timing
        t_end = time.perf_counter() + 0.6
        while time.perf_counter() < t_end:
            step()
        torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    plan.check_device_errors()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * 1e3 / ms_step  # volumes/s, whole job
    info = plan.info
    build_id = plan.lib.build_id()

    # ---- e2e through the reference-facing API with HOST buffers
    del acts, rho, out
    e2e = measure_e2e(wl, args, calc, plan, acts_h, rho_h, dev, dist, world, barrier)
    plan.close()
    calc._plans.clear()
    torch.cuda.empty_cache()

    # ---- the other named configurations, inside the same line (driver-visible): C4 (sharded patients), C5 (slabs + NCCL)
    extras = {}
    if args.workload == "c3" and not args.no_extras:
        xs = max(3, min(args.steps, 10))
        try:
            extras["c4"] = run_c4(args, dev, dist, rank, world, xs)
            torch.cuda.empty_cache()
            for b in ("same", "reference"):
                extras["c5_" + b] = run_c5(args, dev, dist, rank, world, xs, b)
                torch.cuda.empty_cache()
        except Exception as e:  # an extra must never cost the headline line
            extras["error"] = f"{type(e).__name__}: {e}"[:400]

    if rank == 0:
        alg_bytes = 4.0 * (wl["T"] + 1 + (1 if wl["density"] else 0)) * nvox
        kernels = [{"name": n, "ms": round(ms, 4), "hbm_bytes": b, "gbs": round(b / ms / 1e6, 1) if ms > 0 else None,
                    "frac_of_peak": round(b / ms / 1e6 / peak_gbs, 3) if ms > 0 else None} for (n, ms, b) in (acc or [])]
        ksum = sum(k["ms"] for k in kernels) or ms_step
        dom = max(kernels, key=lambda k: k["ms"]) if kernels else None
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        traffic, traffic_src = dram_traffic_for(build_id, args.workload, args.boundary)
        line = {
            "metric": "dose_volumes_per_sec", "value": value, "unit": "volumes/s", "voxels_per_sec": value * nvox,
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": arm_config(wl, args.boundary),
            "impl_config": {"fft_shape": list(info.m), "work_buffer_mb_per_pass": round(info.workspace_bytes / 2e6),
                            "parallelism": f"independent volumes sharded over {world} rank(s), no data-path collective",
                            "library_build_id": build_id},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak_gbs, "unit": "GB/s",
                         "frac": round(achieved / peak_gbs, 4), "traffic": traffic, "traffic_source": traffic_src,
                         "what": "whole conv path (all launches of one volume): algorithmic 4*(T+1+[density]) B/voxel / time per volume",
                         "peak_source": peak_src, "algorithmic_bytes_per_volume": alg_bytes,
                         "implementation_bytes_per_volume": info.hbm_bytes_per_execute,
                         "implementation_gbs": round(info.hbm_bytes_per_execute / (ms_step * 1e-3) / 1e9, 1),
                         "dominant_kernel": dom, "dominant_share_of_step": round(dom["ms"] / ksum, 3) if dom else None},
            "kernels": kernels,
            "e2e": e2e,
            "extras": extras,
            "gpu_launches": int(info.passes) * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, full_affinity)  # the CPU leg may use every host core again
            line["cpu_baseline"] = cpu_baseline_leg(wl, acts_h, times, rho_h)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_leg(wl, acts_h, times, rho_h):
    """The reference's CPU path beside the GPU number, same run, same box, same inputs (float64): the reference's own class
    from oracle/_ref on the FULL volume, 1 warm-up + best of 3, single-threaded by construction;
    plus the same mathematics on every host thread (scipy.fft - not the reference's code)."""
    from oracle import dose_oracle as orc

    nvox = float(np.prod(wl["shape"]))
    maps64 = [a.astype(np.float64) for a in acts_h]
    vox = (float(wl["voxel"]),) * 3
    try:
        calc, k64 = load_reference_calculator(wl)
        kind, what = "reference", "KernelConvolutionCalculator of oracle/_ref (byte-identical copy of the reference)"
        step = lambda: reference_step_real(calc, maps64, times, vox, rho_h)
    except Exception as e:
        sys.stderr.write(f"[bench] reference package unavailable ({e}); timing the oracle port\n")
        k64 = orc.make_kernel(wl["nuclide"], wl["voxel"], wl["kgrid"]).astype(np.float32).astype(np.float64)
        kind, what = "port", "oracle.conv_reference (literal np.fft expression of core/kernel_convolution.py:71-74)"
        step = lambda: reference_step(maps64, k64, rho_h)
    step()  # warm-up
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out = {"value": 1.0 / best, "unit": "volumes/s", "cores": 1, "kind": kind, "seconds_per_volume": round(best, 3),
           "sample": f"{what}, float64, the full {'x'.join(map(str, wl['shape']))} volume, 1 warm-up + best of 3; np.fft is single-threaded",
           "host_cores_available": os.cpu_count()}
    try:
        tv, tc = cpu_threaded_time(wl, acts_h, rho_h, k64, tuple(wl["shape"]))
        out["threaded"] = {"value": tv, "unit": "volumes/s", "cores": tc,
                           "what": "same mathematics via scipy.fft rfftn/irfftn, workers = all host threads, full volume, best of 2 (context only)"}
    except Exception as e:  # pragma: no cover
        out["threaded"] = {"unavailable": str(e)[:200]}
    return out


def run_sharded(args, wl, name):
    """`--workload c4|c5` on their own (the default c3 run carries both in `extras`)."""
    import torch

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group(backend="nccl", device_id=dev)
    xs = max(3, args.steps)
    r = run_c4(args, dev, dist, rank, world, xs) if name == "c4" else run_c5(args, dev, dist, rank, world, xs, args.boundary)
    if rank == 0:
        nvox = float(np.prod(wl["shape"]))
        ms = r.get("ms_per_job", r.get("ms_per_volume"))
        units = wl.get("volumes", 1)
        emit({"metric": "dose_volumes_per_sec", "value": units * 1e3 / ms, "unit": "volumes/s", "voxels_per_sec": units * nvox * 1e3 / ms,
              "n_gpus": world, "steps": xs, "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "f32", "data": "synthetic", "config": {"workload": wl["desc"], "boundary": args.boundary, "l2": "inputs exceed the 126 MB L2"},
              "roofline": r["roofline"], "detail": r, "gpu_launches": r.get("gpu_launches_per_job", r.get("gpu_launches_per_volume", 5)) * xs})
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--boundary", default=None, choices=["reference", "same"],
                    help="default: reference (the reference's circular semantics); c5 defaults to same (zero boundary)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C4 / C5 legs of the default c3 run")
    args = ap.parse_args()
    if args.boundary is None:
        args.boundary = "same" if args.workload == "c5" else "reference"
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    elif args.workload in ("c4", "c5"):
        run_sharded(args, wl, args.workload)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
