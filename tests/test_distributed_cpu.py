"""world_size-2 (and 3) gloo runs of the slab halo exchange on CPU tensors; the local convolution of each
rank runs through the emulated kernels, the stitched result must equal the whole-volume oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, boundary, shape, kshape, q):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from emu_util import EmuConv, emu_lib
        from pyvoxeldosimetry_b200.multi_gpu import exchange_halos, shard_range, slab_geometry

        lib = emu_lib()
        rng = np.random.default_rng(99)
        a = rng.uniform(0, 1e3, shape).astype(np.float32)
        k = rng.uniform(0, 1, kshape).astype(np.float32)
        g = slab_geometry(shape, kshape, boundary, world, rank, lib)
        local = torch.from_numpy(a[g["lo"] : g["hi"]].copy())
        padded = exchange_halos(local, shape[0], boundary, kshape[0])
        assert tuple(padded.shape) == tuple(g["n"])
        kc = g["kcrop"]
        p = EmuConv(lib, g["n"], kc, ex=g["ex"])
        p.set_kernel(k[: kc[0], : kc[1], : kc[2]])
        out = torch.from_numpy(p.execute([padded.numpy()]))
        p.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, (g["lo"], g["hi"], out.numpy()))
        # independent-volume sharding: every volume is owned exactly once
        owned = [None] * world
        dist.all_gather_object(owned, list(shard_range(7, world, rank)))
        if rank == 0:
            full = np.empty(shape, np.float32)
            for lo, hi, arr in gathered:
                full[lo:hi] = arr
            q.put((full, sorted(sum(owned, []))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,boundary,shape,kshape", [
    (2, "same", (12, 6, 10), (5, 3, 3)),
    (2, "reference", (12, 6, 10), (5, 3, 3)),
    (3, "same", (9, 5, 8), (7, 3, 3)),        # halo (3 planes) as wide as a whole neighbour slab
    (3, "reference", (7, 5, 8), (6, 3, 3)),   # halo (5 planes) spans more than one neighbour
])
def test_slab_halo_exchange_gloo(world, boundary, shape, kshape):
    from oracle import dose_oracle as orc

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world * 7 + len(boundary)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, boundary, shape, kshape, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, owned = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(99)
    a = rng.uniform(0, 1e3, shape).astype(np.float32).astype(np.float64)
    k = rng.uniform(0, 1, kshape).astype(np.float32).astype(np.float64)
    ref = orc.conv_reference(a, k) if boundary == "reference" else orc.conv_same(a, k)
    assert orc.rel_err_of_peak(full, ref) <= 1e-4
    assert owned == list(range(7))
