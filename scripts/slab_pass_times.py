"""Per-pass times of ONE rank's local plan of the slab decomposition (1024x1024x800, 51^3), any world size, on one GPU."""
import json, sys
sys.path.insert(0, '.')
import torch
from pyvoxeldosimetry_b200.multi_gpu import SlabConvolver
dev = torch.device('cuda:0')
shape, ks = (1024, 1024, 800), (51, 51, 51)
for world in (8, 4, 2):
    for boundary in ('reference', 'same'):
        sc = SlabConvolver(shape, torch.rand(ks, device=dev), boundary, device=dev, rank=world // 2, world=world)
        sc.padded.copy_(torch.rand(sc.padded.shape, device=dev))
        rho = torch.rand(sc.plan.out_shape, device=dev) + 0.5
        plan = sc.plan
        out = torch.empty(plan.out_shape, device=dev)
        for _ in range(3): plan.execute([sc.padded], None, rho, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): plan.execute([sc.padded], None, rho, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        plan.lib.plan_set_profiling(plan.handle, True)
        acc = None
        for _ in range(5):
            plan.execute([sc.padded], None, rho, out=out)
            pt = plan.lib.plan_get_pass_times(plan.handle)
            acc = acc or [[n, 0.0, by] for n, _, by in pt]
            for i, (_, t, _) in enumerate(pt): acc[i][1] += t / 5
        plan.lib.plan_set_profiling(plan.handle, False)
        print(json.dumps({'world': world, 'boundary': boundary, 'local_fft_shape': list(plan.fft_shape), 'ms': round(ms, 4),
                          'passes': [{'name': n.split(' ')[0], 'ms': round(t, 4), 'GBs': round(by / t / 1e6)} for n, t, by in acc]}))
        del sc, plan, rho, out; torch.cuda.empty_cache()
