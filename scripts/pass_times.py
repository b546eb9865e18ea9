"""Per-pass device times (library CUDA events) of the FFT path for a list of shapes: where does a configuration lose
against C3?  usage: python scripts/pass_times.py [c3 c5 c5slab8 c2 ...]  -> JSON lines"""
import json, sys
sys.path.insert(0, '.')
import torch
from pyvoxeldosimetry_b200.engine import ConvPlan
dev = torch.device('cuda:0')
CASES = {
    'c3': ((512, 512, 400), (51, 51, 51), 'reference'), 'c3same': ((512, 512, 400), (51, 51, 51), 'same'),
    'c2': ((256, 256, 256), (31, 31, 31), 'reference'),
    'c5': ((1024, 1024, 800), (51, 51, 51), 'reference'), 'c5same': ((1024, 1024, 800), (51, 51, 51), 'same'),
    'c5slab8': ((178, 1024, 800), (51, 51, 51), 'reference'), 'c5slab2': ((562, 1024, 800), (51, 51, 51), 'reference'),
}
for name in (sys.argv[1:] or ['c3', 'c5', 'c5same']):
    shape, ks, b = CASES[name]
    plan = ConvPlan(shape, ks, b, dev)
    plan.set_kernel(torch.rand(ks, device=dev))
    a = torch.rand(shape, device=dev); rho = torch.rand(plan.out_shape, device=dev) + 0.5
    out = torch.empty(plan.out_shape, device=dev)
    for _ in range(3): plan.execute([a], None, rho, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): plan.execute([a], None, rho, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    plan.lib.plan_set_profiling(plan.handle, True)
    acc = None
    for _ in range(5):
        plan.execute([a], None, rho, out=out)
        pt = plan.lib.plan_get_pass_times(plan.handle)
        acc = acc or [[n, 0.0, by] for n, _, by in pt]
        for i, (_, t, _) in enumerate(pt): acc[i][1] += t / 5
    plan.lib.plan_set_profiling(plan.handle, False)
    print(json.dumps({'case': name, 'shape': shape, 'fft_shape': list(plan.fft_shape), 'boundary': b, 'ms_per_volume': round(ms, 4),
                      'alg_frac_of_6554': round(12 * a.numel() / ms / 1e6 / 6553.9, 4),
                      'passes': [{'name': n, 'ms': round(t, 4), 'GBs': round(by / t / 1e6, 0)} for n, t, by in acc]}))
    plan.close(); del a, rho, out; torch.cuda.empty_cache()
