"""Where does the halo exchange go?  torchrun, N = 2, a 256 x 1024 x 800 volume: every rank then has exactly the local
problem of one of EIGHT ranks of the 1024-plane volume (128 own planes + 50 halo planes).  Times, per boundary mode:
exchange alone, the forward passes of the own planes alone, and the whole slab convolution overlapped / serialised /
without exchange.  Run with different NCCL_* settings to compare transports."""
import json, os, sys
sys.path.insert(0, '.')
sys.stdout.flush(); _REAL = os.dup(1); os.dup2(2, 1)
import torch, torch.distributed as dist
from pyvoxeldosimetry_b200.multi_gpu import SlabConvolver, _post_exchange

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device(f'cuda:{local}')
dist.init_process_group('nccl', device_id=dev)

def timeit(fn, reps=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 4)

res = {'tag': sys.argv[1] if len(sys.argv) > 1 else 'default', 'env': {k: v for k, v in os.environ.items() if k.startswith('NCCL_')}}
shape = (128 * world, 1024, 800)
for boundary in ('same', 'reference'):
    k = torch.rand((51, 51, 51), device=dev)
    sc = SlabConvolver(shape, k, boundary, device=dev, transport=os.environ.get('PROBE_TRANSPORT', 'auto'))
    sc.interior.copy_(torch.rand(sc.interior.shape, device=dev))
    rho = torch.rand(sc.plan.out_shape, device=dev) + 0.5
    main = torch.cuda.current_stream(dev)
    off, B, L = sc.hplan['own_off'], sc.hi - sc.lo, sc.geom['n'][0]
    def exch():
        if sc.transport == 'peer':
            sc._peer_exchange()
        else:
            for r in _post_exchange(sc.padded, sc.hplan): r.wait()
    def fwd_own():
        sc.plan.lib.conv_forward_planes(sc.plan.handle, [sc.padded.data_ptr()], None, 1.0, off, off + B, main.cuda_stream)
    r = {'transport': sc.transport, 'local_fft_shape': list(sc.plan.fft_shape), 'halo_MB_received': round((L - B) * shape[1] * shape[2] * 4 / 1e6, 1)}
    r['exchange_alone_ms'] = timeit(exch)
    r['forward_own_planes_alone_ms'] = timeit(fwd_own)
    for rs in ((0, 4, 8, 12, 16, 24, 32) if sc.transport == "nccl" else (0,)):
        sc.EXCHANGE_SMS = rs
        r[f'whole_overlapped_reserve{rs}_ms'] = timeit(lambda: sc(density_slab=rho))
    sc.EXCHANGE_SMS = 32
    r['whole_serialised_ms'] = timeit(lambda: sc(density_slab=rho, overlap=False))
    r['whole_no_exchange_ms'] = timeit(lambda: sc(density_slab=rho, exchange=False))
    def split_no_exchange():
        lib, h, p = sc.plan.lib, sc.plan.handle, [sc.padded.data_ptr()]
        lib.conv_forward_planes(h, p, None, 1.0, off, off + B, main.cuda_stream)
        lib.conv_forward_planes(h, p, None, 1.0, 0, off, main.cuda_stream)
        lib.conv_forward_planes(h, p, None, 1.0, off + B, L, main.cuda_stream)
        lib.conv_finish(h, rho.data_ptr(), 0.1, 0.0, sc.out.data_ptr(), main.cuda_stream)
    r['whole_split_no_exchange_ms'] = timeit(split_no_exchange)
    res[boundary] = r
    del sc
    torch.cuda.empty_cache()
if rank == 0:
    os.write(_REAL, (json.dumps(res) + '\n').encode())
dist.barrier(); dist.destroy_process_group()
