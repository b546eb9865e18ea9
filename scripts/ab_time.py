"""A/B timing of library builds: run as  PVDOSE_LIB=path/to/libpvdose_variant.so python scripts/ab_time.py [tag] [cases]
Prints one JSON line per case: ms per volume (CUDA events, 60 steps after warm-up), per-pass times, and the
deviation from a cuFFT (torch.fft) convolution of the same inputs (sanity, of peak)."""
import json
import os
import sys

sys.path.insert(0, '.')
import torch

from pyvoxeldosimetry_b200.engine import ConvPlan

tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get('PVDOSE_LIB', 'default')
want = sys.argv[2].split(',') if len(sys.argv) > 2 else ['c3', 'c3same', 'c2']
CASES = {
    'c3': ((512, 512, 400), (51, 51, 51), 'reference', 1, True),
    'c3same': ((512, 512, 400), (51, 51, 51), 'same', 1, True),
    'c2': ((256, 256, 256), (31, 31, 31), 'reference', 4, False),
    'c2same': ((256, 256, 256), (31, 31, 31), 'same', 4, False),
    'c5slab': ((306, 1024, 800), (51, 51, 51), 'reference', 1, True),
    'c5': ((1024, 1024, 800), (51, 51, 51), 'reference', 1, True),
    'c5same': ((1024, 1024, 800), (51, 51, 51), 'same', 1, True),
    'c5slab8same': ((178, 1024, 800), (51, 51, 51), 'slab8same', 1, True),
}
dev = torch.device('cuda:0')
for name in want:
    shape, ks, boundary, T, den = CASES[name]
    g = torch.Generator(device=dev).manual_seed(3)
    acts = [torch.rand(shape, device=dev, generator=g) for _ in range(T)]
    w = None if T == 1 else [0.5 + 0.25 * i for i in range(T)]
    rho = (torch.rand(shape, device=dev, generator=g) + 0.5) if den else None
    k = torch.rand(ks, device=dev, generator=g)
    if boundary == 'slab8same':  # the local problem of one of 8 ranks (128 planes + 50 halo planes), zero boundary
        from pyvoxeldosimetry_b200._capi import get_lib
        plan = ConvPlan(shape, ks, 'same', dev, ex=dict(m=(180, get_lib().good_fft_size(1049, 1), get_lib().good_fft_size(825, 2)), out_lo=(50, 25, 25), out_n=(128, 1024, 800)))
        rho = rho[:128].contiguous()
    else:
        plan = ConvPlan(shape, ks, boundary, dev)
    plan.set_kernel(k)
    out = torch.empty(plan.out_shape, device=dev)
    for _ in range(5):
        plan.execute(acts, w, rho, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            plan.execute(acts, w, rho, out=out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    plan.lib.plan_set_profiling(plan.handle, True)
    acc = None
    for _ in range(5):
        plan.execute(acts, w, rho, out=out)
        pt = plan.lib.plan_get_pass_times(plan.handle)
        if acc is None:
            acc = [0.0] * len(pt)
        for i, (_, ms, _) in enumerate(pt):
            acc[i] += ms / 5
    plan.lib.plan_set_profiling(plan.handle, False)
    err = None
    if boundary == 'reference' and shape[0] * shape[1] * shape[2] < 3e8:
        kp = torch.zeros(shape, device=dev)
        kp[: ks[0], : ks[1], : ks[2]] = k
        a = acts[0] if T == 1 else sum(w[i] * acts[i] for i in range(T))
        ref = torch.fft.irfftn(torch.fft.rfftn(a) * torch.fft.rfftn(kp), s=shape)
        if rho is not None:
            ref = ref / torch.clamp(rho, min=0.1)
        plan.execute(acts, w, rho, out=out)
        err = float((out - ref).abs().max() / ref.abs().max())
        del ref, kp
    print(json.dumps({"tag": tag, "case": name, "ms": round(best, 4), "passes_ms": [round(x, 4) for x in acc], "err_vs_cufft": err}), flush=True)
    plan.close()
    del acts, rho, out
    torch.cuda.empty_cache()
