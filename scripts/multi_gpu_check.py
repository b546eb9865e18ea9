"""Run under torchrun (one rank per GPU, NCCL): validates the slab decomposition + NCCL halo exchange against a
single-GPU convolution of the whole volume, then times config C5 (1024x1024x800, 51^3) slab-decomposed.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py [--c5]
"""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from pyvoxeldosimetry_b200.engine import ConvPlan
from pyvoxeldosimetry_b200.multi_gpu import SlabConvolver, exchange_halos, shard_range

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device(f'cuda:{local}')
dist.init_process_group('nccl', device_id=dev)
res = {}
for boundary, shape, ks in (('same', (200, 96, 128), (31, 31, 31)), ('reference', (200, 96, 128), (31, 31, 31)), ('same', (67, 40, 64), (8, 5, 7))):
    g = torch.Generator(device='cpu').manual_seed(7)
    a = torch.rand(shape, generator=g)
    k = torch.rand(ks, generator=g)
    sc = SlabConvolver(shape, k, boundary, device=dev, transport=os.environ.get('SLAB_TRANSPORT', 'auto'))
    res['transport'] = sc.transport
    local_in = a[sc.lo:sc.hi].to(dev).contiguous()
    for _ in range(3):   # earlier epochs with other data: a halo left over from them would show in the last result
        sc(torch.rand_like(local_in))
    out = sc(local_in)
    sc.check_device_errors()
    gathered = [None] * world
    dist.all_gather_object(gathered, (sc.lo, sc.hi, out.cpu()))
    if rank == 0:
        full = torch.empty(shape)
        for lo, hi, t in gathered:
            full[lo:hi] = t
        plan = ConvPlan(shape, ks, boundary, dev)
        plan.set_kernel(k)
        ref = plan.execute([a.to(dev)]).cpu()
        err = float((full - ref).abs().max() / ref.abs().max())
        res[f'slab_{boundary}_{"x".join(map(str, shape))}_k{"x".join(map(str, ks))}'] = err
        assert err < 1e-5, (boundary, shape, err)
    dist.barrier()
if '--c5' in sys.argv:
    shape, ks = (1024, 1024, 800), (51, 51, 51)
    k = torch.rand(ks, device=dev)
    sc = SlabConvolver(shape, k, 'same', device=dev)
    local_in = torch.rand((sc.hi - sc.lo,) + shape[1:], device=dev)
    rho = torch.rand((sc.hi - sc.lo,) + shape[1:], device=dev) + 0.5
    for _ in range(2):
        sc(local_in, rho)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    steps = 5
    t_ex = 0.0
    e0.record()
    for _ in range(steps):
        sc(local_in, rho)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # halo exchange alone
    torch.cuda.synchronize(); dist.barrier(); e0.record()
    for _ in range(steps):
        exchange_halos(local_in, shape[0], 'same', ks[0])
    e1.record(); torch.cuda.synchronize()
    mx = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        res['c5'] = {'n_gpus': world, 'ms_per_volume': float(ms), 'volumes_per_s': 1e3 / float(ms), 'voxels_per_s': 1e3 / float(ms) * np.prod(shape),
                     'halo_exchange_ms': float(mx), 'local_fft_shape': list(sc.plan.fft_shape), 'slab_planes': sc.hi - sc.lo}
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
