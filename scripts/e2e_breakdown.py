"""Where does an end-to-end front-door call spend its time?  C3 shape, pageable float32 ndarrays in, ndarray out.
Wall-clock phases with a device synchronise between them (diagnostic only; the bench times the unsplit call).
Output: JSON lines -> profiles/r02_e2e_breakdown.jsonl"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyvoxeldosimetry_b200 import DoseCalculator, engine  # noqa: E402

dev = torch.device("cuda:0")
shape = (512, 512, 400)
rng = np.random.default_rng(1)
a32 = rng.random(shape, dtype=np.float32)
rho = (rng.random(shape, dtype=np.float32) + 0.5).astype(np.float32)
vox = (1.0, 1.0, 1.0)
front = DoseCalculator("Y90", "kernel", {"kernel_grid": (51, 51, 51), "device": "cuda:0", "kernel_resolution": 1.0, "tissue_name": "water"})
kc = front.calculator


def sync():
    torch.cuda.synchronize(dev)


def best(fn, reps=5, warm=2):
    for _ in range(warm):
        r = fn()
    ts = []
    for _ in range(reps):
        sync()
        t0 = time.perf_counter()
        r = fn()
        sync()
        ts.append((time.perf_counter() - t0) * 1e3)
    del r
    return round(float(np.median(ts)), 3)


res = {}
st = engine.HostStager.get(dev)
d_a = torch.empty(shape, dtype=torch.float32, device=dev)
res["stager_upload_419MB_ms"] = best(lambda: st.upload(a32, out=d_a))
pin = torch.empty(shape, dtype=torch.float32).pin_memory()
res["pinned_d2h_419MB_ms"] = best(lambda: pin.copy_(d_a, non_blocking=True))
pin_in = torch.from_numpy(a32).pin_memory()
res["pinned_h2d_419MB_ms"] = best(lambda: d_a.copy_(pin_in, non_blocking=True))
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def both_pinned():
    with torch.cuda.stream(s_in):
        d_a.copy_(pin_in, non_blocking=True)
    with torch.cuda.stream(s_out):
        pin.copy_(d_a, non_blocking=True)


res["pinned_h2d_and_d2h_concurrent_ms"] = best(both_pinned)
d_b = torch.empty(shape, dtype=torch.float32, device=dev)


def stager_up_and_d2h(nch=8):
    bounds = [shape[0] * i // nch for i in range(nch + 1)]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        with torch.cuda.stream(s_in):
            st.upload(rho[lo:hi], out=d_b[lo:hi])
        with torch.cuda.stream(s_out):
            pin[lo:hi].copy_(d_a[lo:hi], non_blocking=True)


for nch in (1, 4, 8, 16):
    res[f"stager_upload_and_pinned_d2h_concurrent_{nch}chunks_ms"] = best(lambda: stager_up_and_d2h(nch))
res["host_call_density_ms"] = best(lambda: kc.calculate_dose_rate(a32, vox, tissue_densities=rho))
res["host_call_no_density_ms"] = best(lambda: kc.calculate_dose_rate(a32, vox))
for nch in (2, 4, 16, 32):
    kc.HOST_PIPELINE_CHUNKS = nch
    res[f"host_call_density_{nch}chunks_ms"] = best(lambda: kc.calculate_dose_rate(a32, vox, tissue_densities=rho))
kc.HOST_PIPELINE_CHUNKS = 8
res["front_door_density_ms"] = best(lambda: front.calculate_dose(activity_maps=[a32], time_points=[2.0], voxel_size=vox, tissue_densities=rho).dose_rate_maps[0])
out = np.empty(shape, dtype=np.float32)
res["host_call_density_out_ndarray_ms"] = best(lambda: kc.calculate_dose_rate(a32, vox, tissue_densities=rho, out=out))
print(json.dumps(res))
