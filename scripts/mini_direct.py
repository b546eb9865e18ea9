import sys
sys.path.insert(0, '.')
import numpy as np, torch
from pyvoxeldosimetry_b200.engine import ConvPlan
dev = torch.device('cuda:0')
shape, ks = (16, 16, 64), (3, 3, 3)
plan = ConvPlan(shape, ks, 'same', dev, 2)
k = torch.rand(ks, device=dev); plan.set_kernel(k)
a = torch.rand(shape, device=dev)
out = plan.execute([a])
torch.cuda.synchronize()
print('ok', out.sum().item())
