"""Multi-GPU partitioning of the dose path (one process per GPU, torch.distributed / NCCL over NVLink).

Two ways the path shards (the reference itself is single-process, core/kernel_convolution.py:48-76):
  1. independent patient volumes / timepoint sets -> `shard_range`: no data-path collective at all;
  2. one very large volume -> contiguous slabs along axis 0 (the slowest memory axis of the C-order arr[x, y, z] layout)
     with a kernel-radius halo exchanged between neighbouring ranks, then an ordinary local convolution on slab+halo
     that keeps only its interior planes (overlap-save).  No distributed FFT / all-to-all is ever needed.

`halo_plan` is the pure index logic (who sends which planes to whom, where they land); `SlabConvolver` owns ONE
preallocated slab-plus-halo buffer per rank, receives straight into views of it and sends straight from views of it
(plane ranges are contiguous memory: no staging copies, no per-call allocation), issues the exchange on a side stream
and runs the plane-local forward passes of its own planes (pvd_conv_forward_planes) while the halo planes are in
flight; the halo planes get the same passes on arrival, then pvd_conv_finish runs the x pass and the inverse passes.
Only index logic and point-to-point exchange live here; the arithmetic is libpvdose (CUDA).
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


def _needs(bounds, r: int, shape0: int, boundary: str, kshape0: int) -> Tuple[int, int]:
    lo, hi = bounds[r]
    if boundary == "reference":
        return lo - (min(kshape0, shape0) - 1), hi
    c0 = kshape0 // 2
    return lo - (kshape0 - 1 - c0), hi + c0


def halo_plan(shape0: int, world: int, boundary: str, kshape0: int) -> List[Dict]:
    """Exchange schedule of every rank (pure index logic).  Entry r:
         own    (lo, hi)        global planes rank r owns
         need   (nlo, nhi)      global plane range of its slab-plus-halo buffer (wraps / sticks out of [0, n0))
         own_off                offset of its own planes inside that buffer
         recvs  [(peer, dst_off, count)]            planes arriving from `peer` land at buffer planes [dst_off, dst_off+count)
         sends  [(peer, src_off, count)]            buffer planes [src_off, ...) (inside the own range) go to `peer`
         copies [(src_off, dst_off, count)]         planes of its own slab it needs again elsewhere (circular wrap onto itself)
       Every needed range is intersected with every rank's owned range, so halos wider than a neighbour's slab work too.
       recvs of rank r from peer p and sends of p to r list the same pieces in the same order (NCCL matches them in order)."""
    if boundary not in ("reference", "same"):
        raise ValueError(f"unknown boundary mode {boundary!r}")
    bounds = slab_bounds(shape0, world)
    wrap = boundary == "reference"
    plans = []
    for r in range(world):
        lo, hi = bounds[r]
        nlo, nhi = _needs(bounds, r, shape0, boundary, kshape0)
        plans.append(dict(own=(lo, hi), need=(nlo, nhi), own_off=lo - nlo, recvs=[], sends=[], copies=[]))
    for r in range(world):
        nlo, nhi = plans[r]["need"]
        lo, hi = bounds[r]
        for dst_off, glo, ghi in _segments(nlo, nhi, shape0, wrap):
            for peer, (plo, phi) in enumerate(bounds):
                a, b = max(glo, plo), min(ghi, phi)
                if a >= b:
                    continue
                d0 = dst_off + (a - glo)
                if peer == r:
                    src = plans[r]["own_off"] + (a - lo)
                    if src != d0:  # the own planes themselves sit at own_off already
                        plans[r]["copies"].append((src, d0, b - a))
                else:
                    plans[r]["recvs"].append((peer, d0, b - a))
                    plans[peer]["sends"].append((r, plans[peer]["own_off"] + (a - plo), b - a))
    return plans


def exchange_halos(local: torch.Tensor, shape0: int, boundary: str, kshape0: int, group=None) -> torch.Tensor:
    """local: this rank's own planes [B, n1, n2] (global planes [lo, hi)).  Returns a NEW slab-plus-halo tensor (the
    allocation-free form is SlabConvolver).  Point-to-point only."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    plan = halo_plan(shape0, world, boundary, kshape0)[rank]
    nlo, nhi = plan["need"]
    out = torch.zeros((nhi - nlo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    off = plan["own_off"]
    out[off : off + local.shape[0]].copy_(local)
    for req in _post_exchange(out, plan, group):
        req.wait()
    for src, dst, cnt in plan["copies"]:
        out[dst : dst + cnt].copy_(out[src : src + cnt])
    return out


def _post_exchange(buf: torch.Tensor, plan: Dict, group=None) -> list:
    """Post every receive (into views of `buf`) and send (from views of `buf`) of `plan` as ONE batch
    (= ncclGroupStart/End around ncclSend/ncclRecv).  Returns the request handles."""
    import torch.distributed as dist

    ops = [dist.P2POp(dist.irecv, buf[d0 : d0 + cnt], peer, group) for peer, d0, cnt in plan["recvs"]]
    ops += [dist.P2POp(dist.isend, buf[s0 : s0 + cnt], peer, group) for peer, s0, cnt in plan["sends"]]
    return dist.batch_isend_irecv(ops) if ops else []


class _StreamWait:
    """Request-like handle of the peer transport: wait() makes `waiter` wait for everything enqueued on `stream` so far."""

    def __init__(self, stream, waiter):
        self.event = torch.cuda.Event()
        self.event.record(stream)
        self.waiter = waiter

    def wait(self):
        self.waiter.wait_event(self.event)


class SlabConvolver:
    """Rank-local half of a slab-decomposed convolution (CUDA).  Usage on every rank:
        sc = SlabConvolver(global_shape, kernel, boundary)          # after init_process_group('nccl')
        sc.interior.copy_(my_planes)   # or hand `local=` to the call; sc.interior is a view of the slab+halo buffer
        dose_slab = sc(density_slab=rho_slab)                       # global planes [sc.lo, sc.hi)
    `rank=` / `world=` given explicitly build the geometry without a process group (single-GPU emulation of every
    rank in turn: tests, 1-GPU runs); halos are then filled with `fill_from_global`."""

    # SMs the forward passes of the own planes leave free while an NCCL exchange is in flight (pvd_plan_reserve_sms)
    EXCHANGE_SMS = 32

    def __init__(self, shape: Sequence[int], kernel, boundary: str = "same", group=None, device=None,
                 rank: Optional[int] = None, world: Optional[int] = None, transport: str = "auto"):
        from .engine import ConvPlan, require_cuda, to_device_f32

        self.group = group
        self.distributed = rank is None
        if self.distributed:
            import torch.distributed as dist

            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:
            self.world, self.rank = int(world), int(rank)
        self.shape = tuple(int(s) for s in shape)
        self.boundary = boundary
        self.device = require_cuda(device)
        kdev = to_device_f32(kernel, self.device)
        self.kshape = tuple(kdev.shape)
        self.geom = slab_geometry(self.shape, self.kshape, boundary, self.world, self.rank)
        self.lo, self.hi = self.geom["lo"], self.geom["hi"]
        self.hplan = halo_plan(self.shape[0], self.world, boundary, self.kshape[0])[self.rank]
        kc = self.geom["kcrop"]
        kdev = kdev[: kc[0], : kc[1], : kc[2]].contiguous()
        self.plan = ConvPlan(self.geom["n"], kc, boundary, self.device, ex=self.geom["ex"])
        self.plan.set_kernel(kdev)
        # ONE slab-plus-halo buffer for the life of the object; halo planes outside the volume ('same' mode at the ends)
        # are zero once and never written again
        self.padded = torch.zeros(self.geom["n"], dtype=torch.float32, device=self.device)
        off, B = self.hplan["own_off"], self.hi - self.lo
        self.interior = self.padded[off : off + B]
        self.out = torch.empty(self.plan.out_shape, dtype=torch.float32, device=self.device)
        self.comm_stream = torch.cuda.Stream(self.device)
        # Halo transport.  "peer": every rank maps its neighbours' slab buffers (CUDA IPC) and PULLS its halo planes
        # with peer-to-peer copies over NVLink - copy engines, no SMs, so the transfer really runs beside the forward
        # passes; NCCL only provides the two tiny device-side barriers around it (all ranks' planes ready / all reads
        # done).  "nccl": send/recv pairs in one group - an SM kernel that shares the SMs with the persistent FFT kernels
        # and is starved by them (profiles/r02_slab_overlap_probe.jsonl); kept for ranks that cannot map each other.
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError("transport must be 'auto', 'peer' or 'nccl'")
        self.transport = "nccl"
        self.peers: Dict[int, torch.Tensor] = {}
        if self.distributed and self.world > 1 and transport in ("auto", "peer"):
            self.transport = "peer" if self._map_peers() else "nccl"
            if transport == "peer" and self.transport != "peer":
                raise RuntimeError("peer transport requested but the ranks cannot map each other's buffers (CUDA IPC)")

    def _map_peers(self) -> bool:
        """Exchange CUDA IPC handles of the slab buffers and open the neighbours' ones.  Collective; every rank gets the
        same answer (the outcome is agreed with an all-reduce)."""
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor

        ok = 1
        # flags[0][p]: "rank p's own planes of epoch e are final" (written by p);  flags[1][p]: "rank p has read what it needs
        # from me in epoch e" (written by p).  uint32 epochs, one slot per rank, in THIS rank's memory.
        self.flags = torch.zeros((2, self.world), dtype=torch.int32, device=self.device)
        self._epoch_word = torch.zeros(1, dtype=torch.int32, device=self.device)  # source of the remote flag copies
        try:
            handle = (reduce_tensor(self.padded), reduce_tensor(self.flags))
        except Exception:
            handle, ok = None, 0
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handle, group=self.group)
        pull_from = sorted({p for p, _, _ in self.hplan["recvs"]})
        push_to = sorted({p for p, _, _ in self.hplan["sends"]})
        self.peer_flags: Dict[int, torch.Tensor] = {}
        try:
            for p in sorted(set(pull_from) | set(push_to)):
                (fn, args), (ffn, fargs) = gathered[p]
                t, f = fn(*args), ffn(*fargs)
                if t.device == self.device or tuple(t.shape)[1:] != tuple(self.padded.shape)[1:]:
                    raise RuntimeError("peer buffer maps onto this rank's own device")
                self.peers[p], self.peer_flags[p] = t, f
                # one framework copy each way: turns on direct peer access between the two devices (without it the
                # driver stages device-to-device copies through the host)
                keep = self.padded[:1].clone()
                self.padded[:1].copy_(t[:1])
                self.padded[:1].copy_(keep)
                f[0, self.rank : self.rank + 1].copy_(self._epoch_word)
                torch.cuda.synchronize(self.device)
            # the stream-memory operations must exist, and a copy into every peer's flag words must be accepted
            st0 = torch.cuda.current_stream(self.device).cuda_stream
            self.plan.lib.stream_write_flag(self._epoch_word.data_ptr(), 0, st0)
            self.plan.lib.stream_wait_flag_geq(self._epoch_word.data_ptr(), 0, st0)
            for p, f in self.peer_flags.items():
                self.plan.lib.copy_async(f[0, self.rank:].data_ptr(), self._epoch_word.data_ptr(), 4, st0)
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            self.peers.clear()
            return False
        self._epoch = 0
        self._pull_from, self._push_to = pull_from, push_to
        # where each piece sits in the peer's buffer: the k-th receive from p pairs with p's k-th send to this rank
        plans = halo_plan(self.shape[0], self.world, self.boundary, self.kshape[0])
        self._pulls = []
        for p in pull_from:
            rec = [(d, c) for (q, d, c) in self.hplan["recvs"] if q == p]
            snd = [(s, c) for (q, s, c) in plans[p]["sends"] if q == self.rank]
            for (d, c), (s0, c2) in zip(rec, snd):
                assert c == c2
                self._pulls.append((p, s0, d, c))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        return True

    def _peer_exchange(self) -> None:
        """On the current (comm) stream, without a single kernel: tell the ranks that pull from me that my planes are
        final, wait for the owners of my halo planes to say the same, pull the planes with peer-to-peer copies (copy
        engines over NVLink), tell the owners I am done, and wait until everybody who reads my planes is done too.
        Waits are stream-memory operations on THIS rank's flag words; a flag in a peer's memory is raised by copying the
        epoch word there (the driver refuses cuStreamWriteValue32 on an IPC-mapped address)."""
        lib = self.plan.lib
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._epoch += 1
        e, r = self._epoch, self.rank
        plane_bytes = self.padded[0].numel() * 4
        lib.stream_write_flag(self._epoch_word.data_ptr(), e, st)
        src = self._epoch_word.data_ptr()
        for p in self._push_to:   # my planes are final (this stream has waited for the compute stream)
            lib.copy_async(self.peer_flags[p][0, r:].data_ptr(), src, 4, st)
        for p in self._pull_from:
            lib.stream_wait_flag_geq(self.flags[0, p:].data_ptr(), e, st)
        for p, s0, d0, cnt in self._pulls:
            lib.copy_async(self.padded[d0:].data_ptr(), self.peers[p][s0:].data_ptr(), cnt * plane_bytes, st)
        for p in self._pull_from:  # done reading p's planes
            lib.copy_async(self.peer_flags[p][1, r:].data_ptr(), src, 4, st)
        for p in self._push_to:    # my planes may be rewritten once every reader is done
            lib.stream_wait_flag_geq(self.flags[1, p:].data_ptr(), e, st)

    # -------------------------------------------------------------- emulation helper (no process group)
    def fill_from_global(self, volume: torch.Tensor) -> None:
        """Fill the whole slab-plus-halo buffer from a global device volume (what the exchange delivers)."""
        nlo, nhi = self.hplan["need"]
        n0 = self.shape[0]
        for dst_off, glo, ghi in _segments(nlo, nhi, n0, self.boundary == "reference"):
            self.padded[dst_off : dst_off + (ghi - glo)].copy_(volume[glo:ghi])

    # -------------------------------------------------------------- one convolution
    def __call__(self, local: Optional[torch.Tensor] = None, density_slab: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None, exchange: bool = True, overlap: bool = True, rho_ref: float = 1.0,
                 rho_min: float = 0.1, rho_cut: float = 0.0, scale: float = 1.0) -> torch.Tensor:
        B = self.hi - self.lo
        if local is not None:
            if tuple(local.shape) != (B,) + self.shape[1:]:
                raise ValueError("local slab has the wrong shape")
            if local.data_ptr() != self.interior.data_ptr():
                self.interior.copy_(local)
        if density_slab is not None and tuple(density_slab.shape) != self.plan.out_shape:
            raise ValueError("density slab has the wrong shape")
        out = self.out if out is None else out
        lib, h = self.plan.lib, self.plan.handle
        main = torch.cuda.current_stream(self.device)
        off, L = self.hplan["own_off"], self.geom["n"][0]
        gain = float(scale) * (float(rho_ref) if density_slab is not None else 1.0)
        ptr = [self.padded.data_ptr()]
        with torch.cuda.device(self.device):
            reqs = []
            if exchange and self.distributed and (self.hplan["recvs"] or self.hplan["sends"]):
                # the exchange runs on its own stream: it needs the interior planes (sends) but nothing else
                self.comm_stream.wait_stream(main)
                with torch.cuda.stream(self.comm_stream):
                    if self.transport == "peer":
                        self._peer_exchange()
                        reqs = [_StreamWait(self.comm_stream, main)]
                    else:
                        reqs = _post_exchange(self.padded, self.hplan, self.group)
                if not overlap:
                    for r in reqs:
                        r.wait()
                    reqs = []
            if exchange:
                for src, dst, cnt in self.hplan["copies"]:
                    self.padded[dst : dst + cnt].copy_(self.padded[src : src + cnt])
            if reqs:
                # own planes first, while the halo planes are in flight ...
                lib.plan_reserve_sms(h, self.EXCHANGE_SMS if self.transport == "nccl" else 0)
                lib.conv_forward_planes(h, ptr, None, gain, off, off + B, main.cuda_stream)
                lib.plan_reserve_sms(h, 0)
                for r in reqs:
                    r.wait()  # stream-level: `main` waits for the NCCL work, the host does not
                # ... then the halo planes below and above
                lib.conv_forward_planes(h, ptr, None, gain, 0, off, main.cuda_stream)
                lib.conv_forward_planes(h, ptr, None, gain, off + B, L, main.cuda_stream)
            else:
                lib.conv_forward_planes(h, ptr, None, gain, 0, L, main.cuda_stream)
            lib.conv_finish(h, None if density_slab is None else density_slab.data_ptr(), float(rho_min), float(rho_cut),
                            out.data_ptr(), main.cuda_stream)
        return out

    def check_device_errors(self) -> None:
        self.plan.check_device_errors()
