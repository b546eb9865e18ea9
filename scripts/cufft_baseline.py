"""Library baseline row (NOT the product): the same circular convolution + density correction written with
torch.fft (cuFFT R2C / C2R, cached kernel spectrum) on the same GPU, timed like bench.py (CUDA events, warm-up,
inputs resident in HBM), and checked against libpvdose's result.  Prints ONE JSON line.

    python scripts/cufft_baseline.py [c3|c2] [steps]
"""
import json
import sys

sys.path.insert(0, '.')
import torch

from pyvoxeldosimetry_b200.engine import ConvPlan

wl = sys.argv[1] if len(sys.argv) > 1 else 'c3'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
shape, ks, T, den = {'c3': ((512, 512, 400), (51, 51, 51), 1, True), 'c2': ((256, 256, 256), (31, 31, 31), 4, False)}[wl]
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(7)
acts = [torch.rand(shape, device=dev, generator=g) for _ in range(T)]
w = torch.tensor([1.0] if T == 1 else [0.5 + 0.25 * i for i in range(T)], dtype=torch.float32)
rho = (torch.rand(shape, device=dev, generator=g) + 0.5) if den else None
k = torch.rand(ks, device=dev, generator=g)

# library path: spectrum cached once (like the plan), per volume: weighted sum, rfftn, multiply, irfftn, density
kpad = torch.zeros(shape, device=dev)
kpad[: ks[0], : ks[1], : ks[2]] = k
spec = torch.fft.rfftn(kpad)
del kpad
wd = w.to(dev)


def lib_step():
    a = acts[0] if T == 1 else sum(wd[i] * acts[i] for i in range(T))
    d = torch.fft.irfftn(torch.fft.rfftn(a) * spec, s=shape)
    if rho is not None:
        d = d / torch.clamp(rho, min=0.1)
    return d


plan = ConvPlan(shape, ks, 'reference', dev)
plan.set_kernel(k)
out = torch.empty(shape, device=dev)
wl_list = None if T == 1 else [float(x) for x in w]


def our_step():
    plan.execute(acts, wl_list, rho, out=out)


def timeit(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


ref = lib_step()
our_step()
torch.cuda.synchronize()
err = float((out - ref).abs().max() / ref.abs().max())
del ref
ms_lib = timeit(lib_step)
ms_ours = timeit(our_step)
print(json.dumps({"workload": wl, "shape": shape, "kernel": ks, "T": T, "density": den, "steps": steps,
                  "cufft_torch_ms": round(ms_lib, 4), "libpvdose_ms": round(ms_ours, 4), "speedup_vs_cufft": round(ms_lib / ms_ours, 3),
                  "max_rel_diff_of_peak": err,
                  "note": "torch.fft.rfftn/irfftn (cuFFT) + elementwise torch ops, kernel spectrum cached; library second opinion, not used by the product"}))
