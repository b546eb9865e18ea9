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


@pytest.mark.parametrize("boundary", ["reference", "same"])
@pytest.mark.parametrize("n0,world,k0", [(12, 2, 5), (9, 3, 7), (7, 3, 6), (64, 8, 51), (16, 1, 5), (10, 4, 64), (1024, 8, 51)])
def test_halo_plan_delivers_every_needed_plane_exactly_once(boundary, n0, world, k0):
    """Pure index logic: executing every rank's schedule on host arrays must reproduce the wrapped / zero-padded plane
    range each rank's local plan expects, with matching send / receive lists per pair."""
    from pyvoxeldosimetry_b200.multi_gpu import halo_plan, slab_bounds

    if n0 < world:
        pytest.skip("more ranks than planes")
    plans = halo_plan(n0, world, boundary, k0)
    bounds = slab_bounds(n0, world)
    vol = np.arange(1, n0 + 1, dtype=np.float64)  # plane id + 1 (0 = zero padding)
    bufs = []
    for r, p in enumerate(plans):
        nlo, nhi = p["need"]
        b = np.zeros(nhi - nlo)
        lo, hi = bounds[r]
        assert p["own"] == (lo, hi) and p["own_off"] == lo - nlo
        b[p["own_off"] : p["own_off"] + hi - lo] = vol[lo:hi]
        bufs.append(b)
    seen = [np.zeros(len(b), dtype=int) for b in bufs]
    for r, p in enumerate(plans):
        seen[r][p["own_off"] : p["own_off"] + bounds[r][1] - bounds[r][0]] += 1
        # pairwise matching: the k-th receive of r from peer is the k-th send of peer to r
        for peer in range(world):
            rec = [(d, c) for (q, d, c) in p["recvs"] if q == peer]
            snd = [(s, c) for (q, s, c) in plans[peer]["sends"] if q == r]
            assert [c for _, c in rec] == [c for _, c in snd]
            for (d, c), (s, _) in zip(rec, snd):
                po = plans[peer]["own_off"]
                assert po <= s and s + c <= po + bounds[peer][1] - bounds[peer][0]  # sends come from the peer's own planes
                bufs[r][d : d + c] = bufs[peer][s : s + c]
                seen[r][d : d + c] += 1
        for s, d, c in p["copies"]:
            bufs[r][d : d + c] = bufs[r][s : s + c]
            seen[r][d : d + c] += 1
    for r, p in enumerate(plans):
        nlo, nhi = p["need"]
        g = np.arange(nlo, nhi)
        if boundary == "reference":
            want = vol[np.mod(g, n0)]
            assert (seen[r] == 1).all()
        else:
            want = np.where((g >= 0) & (g < n0), vol[np.clip(g, 0, n0 - 1)], 0.0)
            assert (seen[r][(g >= 0) & (g < n0)] == 1).all() and (seen[r][(g < 0) | (g >= n0)] == 0).all()
        assert np.array_equal(bufs[r], want)
