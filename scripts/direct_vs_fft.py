"""Crossover measurement: direct TMA-tiled convolution vs FFT path, both boundary modes, 512x512x400."""
import sys
sys.path.insert(0, '.')
import torch
from pyvoxeldosimetry_b200.engine import ConvPlan
dev = torch.device('cuda:0')
shape = (512, 512, 400)
a = torch.rand(shape, device=dev); rho = torch.rand(shape, device=dev) + 0.5
out = torch.empty(shape, device=dev)
import itertools
for K, boundary in itertools.product((3, 5, 7), ('same', 'reference')):
    ks = (K, K, K)
    k = torch.rand(ks, device=dev)
    res = {}
    for name, algo in (('direct', 2), ('fft', 1)):
        plan = ConvPlan(shape, ks, boundary, dev, algo); plan.set_kernel(k)
        for _ in range(3): plan.execute([a], None, rho, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): plan.execute([a], None, rho, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = (ms, out.clone())
        flag = plan.workspace[plan.workspace.numel()-256:plan.workspace.numel()-252].view(torch.int32).item() if algo == 2 else 0
        print(f'K={K} {boundary:9s} {name:6s} fft_shape={plan.fft_shape} {ms:.3f} ms  alg GB/s {12*a.numel()/ms/1e6:.0f}  GFMA/s {a.numel()*K**3/ms/1e6:.0f} errflag={flag}')
        plan.close()
    d = (res['direct'][1] - res['fft'][1]).abs().max().item() / res['fft'][1].abs().max().item()
    print(f'   direct vs fft max rel diff {d:.2e}')
