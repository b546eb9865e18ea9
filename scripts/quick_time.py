import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from pyvoxeldosimetry_b200.engine import ConvPlan
dev = torch.device('cuda:0')
print(torch.cuda.get_device_name(0))
for shape, ks, b in [((512,512,400),(51,51,51),'reference'), ((512,512,400),(51,51,51),'same'), ((256,256,256),(31,31,31),'reference')]:
    plan = ConvPlan(shape, ks, b, dev)
    k = torch.rand(ks, device=dev); plan.set_kernel(k)
    a = torch.rand(shape, device=dev); rho = torch.rand(shape, device=dev) + 0.5
    out = torch.empty(shape, device=dev)
    for _ in range(3): plan.execute([a], None, rho, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): plan.execute([a], None, rho, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(shape, ks, b, 'fft', plan.fft_shape, 'ms/volume %.3f' % ms, 'algGB/s %.0f' % (12*a.numel()/ms/1e6), 'implGB/s %.0f' % (plan.info.hbm_bytes_per_execute/ms/1e6))
    plan.close()
