"""Does running consecutive C2 patients on two streams with two plans (two work buffers) fill the launch / tail bubbles of the
five persistent passes?  One GPU, 16 patients (256^3, 4 time points, 31^3), CUDA events around the whole batch."""
import json, sys
sys.path.insert(0, '.')
import torch
from pyvoxeldosimetry_b200.engine import ConvPlan
dev = torch.device('cuda:0')
shape, ks, T, B = (256, 256, 256), (31, 31, 31), 4, 16
g = torch.Generator(device=dev).manual_seed(3)
k = torch.rand(ks, device=dev, generator=g)
pats = [[torch.rand(shape, device=dev, generator=g) for _ in range(T)] for _ in range(4)]
w = [0.5, 1.0, 1.0, 0.5]
outs = [torch.empty(shape, device=dev) for _ in range(B)]

def run(nstreams):
    plans = [ConvPlan(shape, ks, 'reference', dev) for _ in range(nstreams)]
    for p in plans: p.set_kernel(k)
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    def batch():
        cur = torch.cuda.current_stream(dev)
        for s in streams: s.wait_stream(cur)
        for b in range(B):
            i = b % nstreams
            with torch.cuda.stream(streams[i]):
                plans[i].execute(pats[b % 4], w, None, out=outs[b])
        for s in streams: cur.wait_stream(s)
    for _ in range(3): batch()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); batch(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / B)
    ref = outs[5].clone()
    for p in plans: p.close()
    return best, ref

r = {}
t1, ref1 = run(1)
t2, ref2 = run(2)
t3, _ = run(3)
r = {'ms_per_patient_1_stream': round(t1, 4), 'ms_per_patient_2_streams': round(t2, 4), 'ms_per_patient_3_streams': round(t3, 4),
     'same_result': bool(torch.equal(ref1, ref2))}
print(json.dumps(r))
