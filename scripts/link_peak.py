"""Bare host<->device link ceiling of the box (no library code): pinned cudaMemcpyAsync H2D, D2H and both at once, per
rank and aggregated over ranks (run under torchrun for N > 1), plus what the HOST side of an end-to-end call costs:
pageable->pinned staging copies and float64->float32 conversion with 1..all threads, first-touch of a fresh result
array, cudaHostRegister.  Output: one JSON line (rank 0) -> profiles/r02_link_peak*.json; bench.py's e2e.link_frac is
quoted against `h2d_d2h_concurrent`."""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

# stdout carries exactly one JSON line: NCCL prints its version banner there, so fd 1 points at stderr for the run
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
NB = 512 * 512 * 400 * 4  # one C3 float32 volume
h_in = torch.empty(NB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(NB, dtype=torch.uint8).pin_memory()
d_in = torch.empty(NB, dtype=torch.uint8, device=dev)
d_out = torch.empty(NB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def barrier():
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize(dev)


def timed(fn, reps=5):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / reps
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


res = {"n_gpus": world, "bytes_per_copy": NB}
res["h2d_GBs_aggregate"] = round(world * NB / timed(h2d) / 1e9, 1)
res["d2h_GBs_aggregate"] = round(world * NB / timed(d2h) / 1e9, 1)
t = timed(both)
res["h2d_d2h_concurrent_GBs_aggregate_each_way"] = round(world * NB / t / 1e9, 1)
res["h2d_d2h_concurrent_ms_per_volume_pair"] = round(t * 1e3, 3)

if rank == 0 and world == 1:
    ncpu = len(os.sched_getaffinity(0))
    res["host_threads"] = ncpu
    n = NB // 4
    src32 = np.random.default_rng(0).random(n, dtype=np.float32)
    src64 = src32.astype(np.float64)
    pin32 = h_in.numpy().view(np.float32)

    def par(fn, nthreads):
        bounds = np.linspace(0, n, nthreads + 1).astype(np.int64)
        th = [threading.Thread(target=fn, args=(int(bounds[i]), int(bounds[i + 1]))) for i in range(nthreads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    host = {}
    for nt in sorted({1, 2, 4, 8, ncpu}):
        if nt > ncpu:
            continue
        par(lambda a, b: np.copyto(pin32[a:b], src32[a:b]), nt)
        t32 = min(par(lambda a, b: np.copyto(pin32[a:b], src32[a:b]), nt) for _ in range(3))
        t64 = min(par(lambda a, b: np.copyto(pin32[a:b], src64[a:b], casting="same_kind"), nt) for _ in range(3))
        host[str(nt)] = {"pageable_f32_to_pinned_ms": round(t32 * 1e3, 2), "pageable_f64_to_pinned_f32_ms": round(t64 * 1e3, 2),
                         "f32_copy_GBs_read": round(NB / t32 / 1e9, 1)}
    res["host_staging_numpy_threads"] = host
    # pageable H2D / D2H through the driver's own staging
    pg = torch.from_numpy(src32)
    res["pageable_h2d_ms"] = round(timed(lambda: d_in.view(torch.float32).copy_(pg)) * 1e3, 2)
    outp = torch.empty(n, dtype=torch.float32)
    res["pageable_d2h_ms"] = round(timed(lambda: outp.copy_(d_out.view(torch.float32))) * 1e3, 2)
    # fresh result array: allocation + first touch by all threads
    def fresh():
        o = np.empty(n, dtype=np.float32)
        par(lambda a, b: np.copyto(o[a:b], pin32[a:b]), ncpu)
        return o
    fresh()
    t0 = time.perf_counter()
    for _ in range(3):
        fresh()
    res["fresh_ndarray_fill_all_threads_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
    o = np.empty(n, dtype=np.float32)
    o[:] = 0
    res["reused_ndarray_fill_all_threads_ms"] = round(min(par(lambda a, b: np.copyto(o[a:b], pin32[a:b]), ncpu) for _ in range(3)) * 1e3, 2)
    # pinning the caller's array in place
    rt = torch.cuda.cudart()
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(src32.ctypes.data, NB, 0)
    res["cudaHostRegister_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    res["cudaHostRegister_rc"] = int(rc)
    if int(rc) == 0:
        reg = torch.from_numpy(src32)
        res["registered_h2d_ms"] = round(timed(lambda: d_in.view(torch.float32).copy_(reg, non_blocking=True)) * 1e3, 2)
        t0 = time.perf_counter()
        rt.cudaHostUnregister(src32.ctypes.data)
        res["cudaHostUnregister_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    t0 = time.perf_counter()
    tmp = torch.empty(NB, dtype=torch.uint8).pin_memory()
    res["pin_memory_alloc_419MB_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
if rank == 0:
    os.write(_REAL_STDOUT, (json.dumps(res) + "\n").encode())
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
