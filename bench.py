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
               desc="C5: 1024x1024x800 volume, 51^3 DVK, slabs along axis 0 + kernel-radius halo exchange (NCCL send/recv)"),
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


def bind_to_gpu_numa(local: int) -> str:
    """One rank per GPU: run this rank's host threads (and so its pinned staging buffers, first touch) on the CPUs
    of the GPU's own NUMA node, so that host<->device copies of different ranks do not share one socket's memory
    controllers / inter-socket link.  Best effort; returns a note for stderr."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local)
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
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"rank on GPU {local} ({bdf}) bound to {len(cpus)} local CPUs ({spec})"
        return f"GPU {local} ({bdf}): no usable local CPUs in {spec!r}"
    except Exception as e:  # sysfs layout / permissions differ: keep the default affinity
        return f"NUMA binding skipped: {e}"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
def cpu_reference_time(wl, budget_s: float, acts, rho, kernel64):
    """Literal reference operator (oracle.conv_reference == core/kernel_convolution.py:71-74, float64,
    single-threaded np.fft) on a bounded sub-volume of the workload.  Returns (voxels/s, sample text, seconds)."""
    from oracle import dose_oracle as orc

    full = wl["shape"]
    cands = [full]
    s = list(full)
    for ax in (2, 1, 0, 2, 1, 0):
        s = list(s)
        s[ax] = max(8, s[ax] // 2)
        cands.append(tuple(s))
    # calibrate on the smallest candidate
    small = cands[-1]
    a = acts[0][: small[0], : small[1], : small[2]].astype(np.float64)
    t0 = time.perf_counter()
    orc.conv_reference(a, kernel64)
    rate = a.size / (time.perf_counter() - t0)  # voxels/s, optimistic for the bigger ones
    pick = small
    for c in cands:
        if np.prod(c) / rate * 1.6 <= budget_s:
            pick = c
            break
    subs = [np.ascontiguousarray(m[: pick[0], : pick[1], : pick[2]]).astype(np.float64) for m in acts]
    t0 = time.perf_counter()
    d = reference_step(subs, kernel64, None if rho is None else rho[: pick[0], : pick[1], : pick[2]])
    dt = time.perf_counter() - t0
    what = "literal np.fft.ifftn(fftn(a)*fftn(k,a.shape)).real float64" if len(acts) == 1 else \
        f"literal calculate_absorbed_dose loop: {len(acts)} np.fft convolutions + trapezoid (core/kernel_convolution.py:94-106)"
    return subs[0].size / dt, f"{what} on a {pick[0]}x{pick[1]}x{pick[2]} sub-volume", dt


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
        "config": {"workload": wl["desc"], "boundary": args.boundary, "volumes_per_step_per_gpu": 1,
                   "l2": "inputs (>=419 MB per volume for c3) exceed the 126 MB L2",
                   "parallelism": "one host process, one thread (the reference has no parallelism)"},
        "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(), "finite": bool(np.isfinite(d).all())},
        "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
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
    sys.stderr.write(bind_to_gpu_numa(local) + "\n")
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
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * 1e3 / ms_step  # volumes/s, whole job

    # ---- e2e through the public calculator API with pinned host buffers
    pin_acts = [torch.from_numpy(a).pin_memory() for a in acts_h]
    pin_rho = None if rho_h is None else torch.from_numpy(rho_h).pin_memory()
    pin_out = torch.empty(plan.out_shape, dtype=torch.float32).pin_memory()
    vox = (wl["voxel"],) * 3

    def e2e_step():
        if wl["T"] == 1:
            return calc.calculate_dose_rate(pin_acts[0], vox, tissue_densities=pin_rho, out=pin_out)
        return calc.calculate_absorbed_dose(pin_acts, times, vox, tissue_densities=pin_rho, out=pin_out)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_variants = {}
    h2d_den_bytes = pin_rho.numel() * 4 if pin_rho is not None else 0
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    dt_single = (time.perf_counter() - t0) / e2e_steps  # one blocking call per volume
    dt, e2e_api = dt_single, "KernelConvolutionCalculator.calculate_dose_rate(host ndarray, tissue_densities=host ndarray, out=pinned)"
    if wl["T"] == 1:
        # pipelined batch call: every volume still pays its own H2D (activity + density) and D2H, but copies in
        # both directions and the convolution overlap across consecutive volumes (three streams, double buffers)
        pin_out2 = torch.empty(plan.out_shape, dtype=torch.float32).pin_memory()
        nb = max(8, e2e_steps)
        batch_acts = [pin_acts[0]] * nb
        batch_den = None if pin_rho is None else [pin_rho] * nb
        batch_outs = [pin_out if i % 2 == 0 else pin_out2 for i in range(nb)]
        calc.calculate_dose_rate_batch(batch_acts[:3], vox, None if batch_den is None else batch_den[:3], batch_outs[:3])
        barrier()
        t0 = time.perf_counter()
        calc.calculate_dose_rate_batch(batch_acts, vox, batch_den, batch_outs)
        torch.cuda.synchronize(dev)
        dt_batch = (time.perf_counter() - t0) / nb
        if dt_batch < dt:
            dt, e2e_api = dt_batch, f"KernelConvolutionCalculator.calculate_dose_rate_batch({nb} host volumes + {nb} host density volumes -> {nb} host dose maps), pipelined H2D/compute/D2H"
        e2e_variants["batch_float_density_ms"] = dt_batch * 1e3
        if pin_rho is not None:
            # the CT as scanners store it: int16 Hounsfield units (2 bytes per voxel over the link), turned into the same
            # densities on the device (HU -1000 / -700 / 32 / 350 <-> 0.00129 / 0.26 / 1.04 / 1.42 g/cm3 through tissue.HU_KNOTS)
            hu_h = np.full(wl["shape"], -1000, dtype=np.int16)
            hu_h[rho_h > 0.2] = -700
            hu_h[rho_h > 1.0] = 32
            hu_h[rho_h > 1.4] = 350
            pin_hu = torch.from_numpy(hu_h).pin_memory()
            calc.calculate_dose_rate_batch(batch_acts[:3], vox, None, batch_outs[:3], ct_hu=[pin_hu] * 3)
            barrier()
            t0 = time.perf_counter()
            calc.calculate_dose_rate_batch(batch_acts, vox, None, batch_outs, ct_hu=[pin_hu] * nb)
            torch.cuda.synchronize(dev)
            dt_ct = (time.perf_counter() - t0) / nb
            e2e_variants["batch_int16_ct_ms"] = dt_ct * 1e3
            if dt_ct < dt:
                dt, e2e_api = dt_ct, f"KernelConvolutionCalculator.calculate_dose_rate_batch({nb} host activity volumes (float32) + {nb} host CT volumes (int16 HU, density derived on the device) -> {nb} host dose maps), pipelined H2D/compute/D2H"
                h2d_den_bytes = pin_hu.numel() * 2
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world / float(t.item())
    h2d = int(sum(a.numel() for a in pin_acts) * 4 + h2d_den_bytes)
    d2h = int(pin_out.numel() * 4)

    if rank == 0:
        alg_bytes = 4.0 * (wl["T"] + 1 + (1 if wl["density"] else 0)) * nvox
        info = plan.info
        kernels = [{"name": n, "ms": round(ms, 4), "hbm_bytes": b, "gbs": round(b / ms / 1e6, 1) if ms > 0 else None,
                    "frac_of_peak": round(b / ms / 1e6 / peak_gbs, 3) if ms > 0 else None} for (n, ms, b) in (acc or [])]
        ksum = sum(k["ms"] for k in kernels) or ms_step
        dom = max(kernels, key=lambda k: k["ms"]) if kernels else None
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        traffic = None  # dram__bytes_read+write of the whole path from the committed ncu --set full capture (same workload only)
        tpath = os.path.join(REPO, "profiles", "r01h_dram_traffic_c3.json")
        if args.workload == "c3" and args.boundary == "reference" and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("_whole_path_dram_bytes")
        line = {
            "metric": "dose_volumes_per_sec", "value": value, "unit": "volumes/s", "voxels_per_sec": value * nvox,
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "boundary": args.boundary, "fft_shape": list(plan.fft_shape),
                       "volumes_per_step_per_gpu": 1, "l2": "inputs (>=419 MB per volume for c3) exceed the 126 MB L2",
                       "parallelism": f"independent volumes sharded over {world} rank(s), no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak_gbs, "unit": "GB/s",
                         "frac": round(achieved / peak_gbs, 4), "traffic": traffic,
                         "what": "whole conv path (all launches of one volume): algorithmic 4*(T+1+[density]) B/voxel / time per volume",
                         "peak_source": peak_src, "algorithmic_bytes_per_volume": alg_bytes,
                         "implementation_bytes_per_volume": info.hbm_bytes_per_execute,
                         "implementation_gbs": round(info.hbm_bytes_per_execute / (ms_step * 1e-3) / 1e9, 1),
                         "dominant_kernel": dom, "dominant_share_of_step": round(dom["ms"] / ksum, 3) if dom else None},
            "kernels": kernels,
            "e2e": {"value": e2e_value, "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": dt * 1e3, "single_call_ms": dt_single * 1e3, "api": e2e_api, "variants_ms": e2e_variants},
            "gpu_launches": int(info.passes) * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import dose_oracle as orc

            os.sched_setaffinity(0, full_affinity)  # the CPU leg may use every host core again

            k64 = calc.kernel.astype(np.float32).astype(np.float64)
            rate, sample, secs = cpu_reference_time(wl, 20.0, acts_h, rho_h, k64)
            line["cpu_baseline"] = {"value": rate / nvox, "unit": "volumes/s", "cores": 1, "kind": "port",
                                    "sample": sample + f" ({secs:.1f} s); np.fft is single-threaded", "host_cores_available": os.cpu_count()}
            try:
                pick = tuple(int(x) for x in sample.split(" on a ")[1].split(" ")[0].split("x"))
                tv, tc = cpu_threaded_time(wl, acts_h, rho_h, k64, pick)
                line["cpu_baseline"]["threaded"] = {"value": tv, "unit": "volumes/s", "cores": tc,
                                                    "what": "same mathematics via scipy.fft rfftn/irfftn, workers = all host threads, same sample"}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"]["threaded"] = {"unavailable": str(e)[:200]}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_sharded(args, wl, name):
    """C4 / C5: a fixed job shared by the ranks (strong scaling).  C4 shards independent volumes (no collective);
    C5 splits one volume into slabs and exchanges kernel-radius halos between neighbours with NCCL send/recv."""
    import torch

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group(backend="nccl", rank=0, world_size=1, device_id=dev)
    from pyvoxeldosimetry_b200 import KernelConvolutionCalculator
    from pyvoxeldosimetry_b200.engine import ConvPlan
    from pyvoxeldosimetry_b200.multi_gpu import SlabConvolver, shard_range

    boundary = args.boundary
    calc = KernelConvolutionCalculator(wl["nuclide"], "water", wl["voxel"], config={"kernel_grid": wl["kgrid"], "boundary": boundary, "device": str(dev)})
    kdev = calc._kernel_dev
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    if name == "c4":
        from pyvoxeldosimetry_b200.core import trapezoid_weights

        mine = shard_range(wl["volumes"], world, rank)
        plan = ConvPlan(wl["shape"], wl["kgrid"], boundary, dev)
        plan.set_kernel(kdev)
        sets = [[torch.rand(wl["shape"], device=dev, generator=g) for _ in range(wl["T"])] for _ in range(2)]  # two patients' buffers, alternated
        w = trapezoid_weights([4.0, 24.0, 96.0, 168.0], 3600.0)
        out = torch.empty(plan.out_shape, device=dev)

        def step():
            for v in mine:
                plan.execute(sets[v & 1], w, None, out=out)
        units, extra, launches = wl["volumes"], {"volumes_per_rank": len(mine), "fft_shape": list(plan.fft_shape)}, 5 * len(mine)
    else:
        sc = SlabConvolver(wl["shape"], kdev, boundary, device=dev)
        local_in = torch.rand((sc.hi - sc.lo,) + tuple(wl["shape"][1:]), device=dev, generator=g)
        rho = torch.rand((sc.hi - sc.lo,) + tuple(wl["shape"][1:]), device=dev, generator=g) + 0.5

        def step():
            sc(local_in, rho)
        units, launches = 1, 5
        extra = {"slab_planes": sc.hi - sc.lo, "halo_planes": sc.geom["n"][0] - (sc.hi - sc.lo), "local_fft_shape": list(sc.plan.fft_shape)}
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize(dev)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    if rank == 0:
        nvox = float(np.prod(wl["shape"]))
        value = units * 1e3 / ms_step
        alg = 4.0 * (wl["T"] + 1 + (1 if wl["density"] else 0)) * nvox * units
        peak_gbs, peak_src = peaks()
        emit({
            "metric": "dose_volumes_per_sec", "value": value, "unit": "volumes/s", "voxels_per_sec": value * nvox, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": wl["desc"], "boundary": boundary, "l2": "inputs exceed the 126 MB L2"}, **extra),
            "roofline": {"bound": "hbm", "achieved": round(alg / (ms_step * 1e-3) / 1e9, 1), "peak": peak_gbs * world, "unit": "GB/s",
                         "frac": round(alg / (ms_step * 1e-3) / 1e9 / (peak_gbs * world), 4), "traffic": None, "peak_source": peak_src + f" x {world} GPUs"},
            "gpu_launches": launches * args.steps,
        })
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
