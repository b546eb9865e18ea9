"""Multi-GPU partitioning of the dose path (one process per GPU, torch.distributed / NCCL over NVLink).

Two ways the path shards:
  1. independent patient volumes / timepoint sets -> `shard_range`: no data-path collective at all;
  2. one very large volume -> contiguous slabs along axis 0 (the slowest memory axis of the C-order
     arr[x, y, z] layout) with a kernel-radius halo exchanged between neighbouring ranks
     (`exchange_halos`: batched isend/irecv = ncclSend/ncclRecv pairs in one group), then an ordinary
     local convolution on slab+halo that keeps only its interior planes (overlap-save).  No distributed
     FFT / all-to-all is ever needed.
Only index logic and point-to-point exchange live here; the arithmetic is ConvPlan (CUDA).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch


def shard_range(n_items: int, world: int, rank: int) -> range:
    """Contiguous share of `n_items` independent volumes for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def slab_bounds(n0: int, world: int) -> List[Tuple[int, int]]:
    return [(r.start, r.stop) for r in (shard_range(n0, world, k) for k in range(world))]


def good_size_py(n: int, lib=None, axis: int = 0) -> int:
    if lib is not None:
        return lib.good_fft_size(n, axis)
    from ._capi import get_lib

    return get_lib().good_fft_size(n, axis)


def slab_geometry(shape: Sequence[int], kshape: Sequence[int], boundary: str, world: int, rank: int, lib=None) -> Dict:
    """Local problem of `rank`: which global planes it needs, and the expert plan geometry.

    Returns dict(lo, hi, need_lo, need_hi, kcrop, n, ex=dict(m, out_lo, out_n)) where the local input
    holds global planes [need_lo, need_hi) (indices wrap in reference mode, are zero outside
    [0, n0) in same mode) and the local output is global planes [lo, hi).
    """
    n0, n1, n2 = (int(s) for s in shape)
    lo, hi = slab_bounds(n0, world)[rank]
    B = hi - lo
    if B < 1:
        raise ValueError(f"rank {rank} would own no planes: {n0} planes over {world} ranks")
    if boundary == "reference":
        ke = tuple(min(int(k), n) for k, n in zip(kshape, (n0, n1, n2)))  # np.fft.fftn(kernel, s=shape) crop
        h = ke[0] - 1
        L = B + h
        m0 = good_size_py(L, lib)
        return dict(lo=lo, hi=hi, need_lo=lo - h, need_hi=hi, kcrop=ke, n=(L, n1, n2),
                    ex=dict(m=(m0, n1, n2), out_lo=(h, 0, 0), out_n=(B, n1, n2)))
    if boundary == "same":
        k0, k1, k2 = (int(k) for k in kshape)
        c = (k0 // 2, k1 // 2, k2 // 2)
        # d[g] = sum_t k[t] a[g + c0 - t]: K0-1-c0 planes below the slab, c0 planes above it
        dn = k0 - 1 - c[0]
        L = B + k0 - 1
        m = [good_size_py(L, lib)]
        for axis, (n, k, cc) in enumerate(((n1, k1, c[1]), (n2, k2, c[2])), start=1):
            m.append(good_size_py(max(n + k - 1 - cc, k, n + cc), lib, axis))
        return dict(lo=lo, hi=hi, need_lo=lo - dn, need_hi=hi + c[0], kcrop=(k0, k1, k2), n=(L, n1, n2),
                    ex=dict(m=tuple(m), out_lo=(k0 - 1, c[1], c[2]), out_n=(B, n1, n2)))
    raise ValueError(f"unknown boundary mode {boundary!r}")


def _segments(need_lo: int, need_hi: int, n0: int, wrap: bool) -> List[Tuple[int, int, int]]:
    """Split the needed global plane range into (dst_offset, global_lo, global_hi) pieces inside [0, n0)."""
    out = []
    g = need_lo
    while g < need_hi:
        if wrap:
            base = (g // n0) * n0  # floor division also for negatives
            seg_hi = min(need_hi, base + n0)
            out.append((g - need_lo, g - base, seg_hi - base))
            g = seg_hi
        else:
            if g < 0:
                g = min(0, need_hi)
                continue
            if g >= n0:
                break
            seg_hi = min(need_hi, n0)
            out.append((g - need_lo, g, seg_hi))
            g = seg_hi
    return out


def exchange_halos(local: torch.Tensor, shape0: int, boundary: str, kshape0: int, group=None) -> torch.Tensor:
    """local: this rank's own planes [B, n1, n2] (global planes [lo, hi)).  Returns the slab-plus-halo
    tensor the local plan consumes.  Point-to-point only: every needed plane range is intersected with
    every peer's owned range, so halos wider than a neighbour's slab are handled too."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = slab_bounds(shape0, world)
    wrap = boundary == "reference"

    def needs(r):
        lo, hi = bounds[r]
        if wrap:
            h = min(kshape0, shape0) - 1
            return lo - h, hi
        c0 = kshape0 // 2
        return lo - (kshape0 - 1 - c0), hi + c0

    my_lo, my_hi = bounds[rank]
    nlo, nhi = needs(rank)
    out = torch.zeros((nhi - nlo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    ops, recvs = [], []
    # receives: pieces of my needed range owned by peers (or myself)
    for dst_off, glo, ghi in _segments(nlo, nhi, shape0, wrap):
        for peer, (plo, phi) in enumerate(bounds):
            a, b = max(glo, plo), min(ghi, phi)
            if a >= b:
                continue
            d0 = dst_off + (a - glo)
            if peer == rank:
                out[d0 : d0 + (b - a)].copy_(local[a - my_lo : b - my_lo])
            else:
                buf = torch.empty((b - a,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
                recvs.append((buf, d0))
                ops.append(dist.P2POp(dist.irecv, buf, peer, group))
    # sends: pieces of every peer's needed range that I own
    keep = []
    for peer in range(world):
        if peer == rank:
            continue
        plo_n, phi_n = needs(peer)
        for _, glo, ghi in _segments(plo_n, phi_n, shape0, wrap):
            a, b = max(glo, my_lo), min(ghi, my_hi)
            if a >= b:
                continue
            piece = local[a - my_lo : b - my_lo].contiguous()
            keep.append(piece)
            ops.append(dist.P2POp(dist.isend, piece, peer, group))
    if ops:
        # order recvs/sends deterministically per pair: batch_isend_irecv groups them (ncclGroupStart/End)
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, d0 in recvs:
        out[d0 : d0 + buf.shape[0]].copy_(buf)
    return out


class SlabConvolver:
    """Rank-local half of a slab-decomposed convolution (CUDA).  Usage on every rank:
        sc = SlabConvolver(global_shape, kernel, boundary)          # after init_process_group('nccl')
        dose_slab = sc(local_activity_slab, density_slab=None)      # global planes [sc.lo, sc.hi)
    """

    def __init__(self, shape: Sequence[int], kernel, boundary: str = "same", group=None, device=None):
        import torch.distributed as dist

        from .engine import ConvPlan, require_cuda, to_device_f32

        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.shape = tuple(int(s) for s in shape)
        self.boundary = boundary
        self.device = require_cuda(device)
        kdev = to_device_f32(kernel, self.device)
        self.kshape = tuple(kdev.shape)
        self.geom = slab_geometry(self.shape, self.kshape, boundary, self.world, self.rank)
        self.lo, self.hi = self.geom["lo"], self.geom["hi"]
        kc = self.geom["kcrop"]
        kdev = kdev[: kc[0], : kc[1], : kc[2]].contiguous()
        self.plan = ConvPlan(self.geom["n"], kc, boundary, self.device, ex=self.geom["ex"])
        self.plan.set_kernel(kdev)

    def __call__(self, local: torch.Tensor, density_slab: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
        if tuple(local.shape) != (self.hi - self.lo,) + self.shape[1:]:
            raise ValueError("local slab has the wrong shape")
        padded = exchange_halos(local, self.shape[0], self.boundary, self.kshape[0], self.group)
        return self.plan.execute([padded], None, density_slab, **kw)
