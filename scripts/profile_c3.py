"""Minimal driver for ncu: a few executes of one plan.  argv: boundary iters shape kernel T density(0/1); default = the C3 plan
(512x512x400, 51^3, density, reference mode)."""
import sys
sys.path.insert(0, '.')
import torch
from pyvoxeldosimetry_b200.engine import ConvPlan
boundary = sys.argv[1] if len(sys.argv) > 1 else 'reference'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shape = tuple(int(x) for x in sys.argv[3].split('x')) if len(sys.argv) > 3 else (512, 512, 400)
ks = tuple(int(x) for x in sys.argv[4].split('x')) if len(sys.argv) > 4 else (51, 51, 51)
T = int(sys.argv[5]) if len(sys.argv) > 5 else 1
den = (sys.argv[6] != '0') if len(sys.argv) > 6 else True
dev = torch.device('cuda:0')
plan = ConvPlan(shape, ks, boundary, dev)
plan.set_kernel(torch.rand(ks, device=dev))
acts = [torch.rand(shape, device=dev) for _ in range(T)]
w = None if T == 1 else [0.5 + 0.25 * t for t in range(T)]
rho = (torch.rand(shape, device=dev) + 0.5) if den else None
out = torch.empty(plan.out_shape, device=dev)
for _ in range(iters):
    plan.execute(acts, w, rho, out=out)
torch.cuda.synchronize()
print('done', plan.fft_shape)
