"""How fast do the five passes run when their input is L2-resident?  (sizing experiment for pass fusion)

The C3 plan streams a 436 MB work buffer through HBM in every pass.  A thin plan (34 x 512 x 400: 29 MB work buffer,
same y / z kernels, 2.99 persistent waves in the y passes) leaves every intermediate in the 126 MB L2, so its
per-tile times are what a fused (L2-blocked) schedule could reach.  Prints one JSON line per plan with per-pass
times, tiles and microseconds per wave of resident CTAs."""
import json
import sys

sys.path.insert(0, '.')
import torch

from pyvoxeldosimetry_b200.engine import ConvPlan

dev = torch.device('cuda:0')
SMS = torch.cuda.get_device_properties(0).multi_processor_count
CASES = [((512, 512, 400), 'c3'), ((34, 512, 400), 'thin34'), ((37, 512, 400), 'thin37'), ((74, 512, 400), 'thin74')]
for shape, name in CASES:
    ks = (51, 51, 51) if shape[0] >= 51 else (min(shape[0], 51), 51, 51)
    g = torch.Generator(device=dev).manual_seed(3)
    a = torch.rand(shape, device=dev, generator=g)
    rho = torch.rand(shape, device=dev, generator=g) + 0.5
    k = torch.rand(ks, device=dev, generator=g)
    plan = ConvPlan(shape, ks, 'reference', dev)
    plan.set_kernel(k)
    out = torch.empty(plan.out_shape, device=dev)
    for _ in range(5):
        plan.execute([a], None, rho, out=out)
    torch.cuda.synchronize()
    plan.lib.plan_set_profiling(plan.handle, True)
    acc, names, byts = None, None, None
    R = 20
    for _ in range(R):
        plan.execute([a], None, rho, out=out)
        pt = plan.lib.plan_get_pass_times(plan.handle)
        if acc is None:
            acc = [0.0] * len(pt)
            names = [p[0] for p in pt]
            byts = [p[2] for p in pt]
        for i, (_, ms, _) in enumerate(pt):
            acc[i] += ms / R
    plan.lib.plan_set_profiling(plan.handle, False)
    n0, n1, n2 = shape
    row_tiles = n0 * n1 // 32
    col_tiles = n0 * ((n2 // 2 + 1 + 15) // 16)
    tiles = [row_tiles, col_tiles, None, col_tiles, row_tiles]
    ctas = [2 * SMS, SMS, None, SMS, 2 * SMS]
    rec = {"case": name, "shape": shape, "work_buffer_MB": round(n0 * n1 * 208 * 8 / 1e6, 1), "passes": []}
    for i, nm in enumerate(names):
        d = {"name": nm.split(' ')[0], "ms": round(acc[i], 4), "GBs": round(byts[i] / acc[i] / 1e6, 0)}
        if i < len(tiles) and tiles[i]:
            waves = tiles[i] / ctas[i]
            d["tiles"] = tiles[i]
            d["waves"] = round(waves, 2)
            d["us_per_wave"] = round(acc[i] * 1e3 / -(-tiles[i] // ctas[i]), 3)
        rec["passes"].append(d)
    print(json.dumps(rec), flush=True)
    plan.close()
    del a, rho, out
    torch.cuda.empty_cache()
