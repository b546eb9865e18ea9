"""Small-size tour of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python scripts/sanitize.py
FFT path on the generic engine and on menu lengths (TMA-staged persistent kernels, programmatic dependent launch), both boundary
modes, time-weighted first pass, density epilogue, cubic + generic direct convolution, split (slab) form, the next-row kernels."""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from pyvoxeldosimetry_b200 import engine
from pyvoxeldosimetry_b200.engine import ConvPlan
from pyvoxeldosimetry_b200.core.utils import interpolate_timepoints, calculate_dvh
from pyvoxeldosimetry_b200.tissue.density import HU_KNOTS

dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(7)
rnd = lambda *s: torch.rand(s, device=dev, generator=g)

def conv(shape, ks, boundary, T=1, den=True, algo=None):
    kw = {} if algo is None else {'algo': algo}
    plan = ConvPlan(shape, ks, boundary, dev, **kw)
    plan.set_kernel(rnd(*ks))
    acts = [rnd(*shape) for _ in range(T)]
    w = None if T == 1 else [0.5 + 0.25 * t for t in range(T)]
    out = plan.execute(acts, w, (rnd(*shape) + 0.5) if den else None)
    plan.check_device_errors()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    plan.close()
    print('conv', shape, ks, boundary, T, den, algo, 'ok', flush=True)

conv((40, 36, 44), (7, 5, 9), 'reference')             # generic engine, any length
conv((40, 36, 44), (7, 5, 9), 'same', T=3)
conv((64, 256, 400), (9, 9, 9), 'reference')            # menu: rows 400 (TMA rows), y 256 (TMA columns), x generic
conv((256, 256, 256), (5, 5, 5), 'reference', T=4, den=False, algo=1)   # C2 kernels
conv((288 - 8, 288 - 8, 288 - 8), (9, 9, 9), 'same', algo=1)            # 288 menu ('same' mode of 280^3): cropped TMA rows, x walk
conv((180, 64, 400), (5, 5, 5), 'reference', algo=1)    # 180-point x walk (8-rank slab length)
conv((96, 80, 72), (5, 5, 5), 'same')                   # cubic direct (AUTO)
conv((96, 80, 72), (5, 5, 5), 'reference')              # cubic direct, circular
conv((50, 40, 72), (3, 5, 7), 'same', algo=2)           # generic direct
# next-row kernels
times = [4.0, 24.0, 96.0, 168.0]
vols = [rnd(40, 33, 29) * 1e3 * float(np.exp(-0.005 * t)) for t in times]
engine.monoexp_fit(vols, times, None, float(np.log(2) / 161.52), 16152.0, want_params=True)
interpolate_timepoints(times, vols, [0.0, 10.0, 50.0, 100.0, 170.0], 'cubic')
interpolate_timepoints(times, vols, list(np.linspace(1, 160, 9)), 'linear')
hu = rnd(40, 33, 29) * 2800 - 1000
hu[3:5, 4:6, 5:7] = 3000.0
engine.ct_prepare(hu, 2000.0, HU_KNOTS, [(-1000, -900), (-900, -500), (-100, 100), (300, 3000), (-10, 10)], want_corrected=True)
dose = rnd(40, 33, 29) * 50
calculate_dvh(dose, (rnd(40, 33, 29) > 0.5), bins=100)
torch.cuda.synchronize()
print('sanitize tour done')
