"""GPU tests of the z-slab decomposition that need ONE GPU: every rank's local plan is run in
turn on the same device (`SlabConvolver(rank=, world=)`, halos filled from the global volume exactly as the NCCL exchange
delivers them), the slabs are stitched and compared with the whole-volume plan and with the float64 oracle.  The
exchange itself is covered by tests/test_distributed_cpu.py (gloo, schedule logic), by the two-process test at the end of
this file where two GPUs are visible, and by bench.py's C5 leg, which checks the stitched slab result against a single-GPU
run on every N > 1 launch."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dose_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def test_split_form_is_bit_identical_to_execute():
    """pvd_conv_forward_planes over any partition of the planes + pvd_conv_finish == pvd_conv_execute."""
    from pyvoxeldosimetry_b200.engine import ConvPlan

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(21)
    for shape, ks, boundary in (((96, 64, 128), (9, 9, 9), "reference"), ((70, 40, 100), (15, 7, 11), "same"), ((192, 192, 256), (7, 7, 7), "reference")):
        a = torch.rand(shape, device=dev, generator=g) * 1e3
        rho = torch.rand(shape, device=dev, generator=g) + 0.3
        plan = ConvPlan(shape, ks, boundary, dev)
        plan.set_kernel(torch.rand(ks, device=dev, generator=g))
        want = plan.execute([a], None, rho, rho_ref=1.1, rho_min=0.4, scale=2.0).clone()
        got = torch.full_like(want, float("nan"))
        n0 = shape[0]
        stream = torch.cuda.current_stream(dev).cuda_stream
        for lo, hi in ((n0 // 3, 2 * n0 // 3), (0, n0 // 3), (2 * n0 // 3, n0)):  # any order
            plan.lib.conv_forward_planes(plan.handle, [a.data_ptr()], None, 2.0 * 1.1, lo, hi, stream)
        plan.lib.conv_finish(plan.handle, rho.data_ptr(), 0.4, 0.0, got.data_ptr(), stream)
        plan.check_device_errors()
        assert torch.equal(got, want), (shape, boundary)
        plan.close()


@pytest.mark.parametrize("boundary", ["same", "reference"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_emulated_ranks_stitch_to_the_whole_volume_and_the_oracle(world, boundary):
    from pyvoxeldosimetry_b200.engine import ConvPlan
    from pyvoxeldosimetry_b200.multi_gpu import SlabConvolver

    dev = torch.device("cuda:0")
    shape, ks = (200, 96, 128), (31, 31, 31)
    rng = np.random.default_rng(7)
    a = rng.uniform(0, 1e3, shape).astype(np.float32)
    a[90:110, 40:60, 50:80] = 2e6
    k = rng.uniform(0, 1, ks).astype(np.float32)
    rho = rng.choice([0.26, 1.04, 1.42], size=shape).astype(np.float32)
    ad, rd = torch.from_numpy(a).to(dev), torch.from_numpy(rho).to(dev)
    full = torch.empty(shape, device=dev)
    for rank in range(world):
        sc = SlabConvolver(shape, k, boundary, device=dev, rank=rank, world=world)
        sc.fill_from_global(ad)
        full[sc.lo : sc.hi] = sc(density_slab=rd[sc.lo : sc.hi].contiguous(), exchange=False)
        sc.check_device_errors()
        sc.plan.close()
    plan = ConvPlan(shape, ks, boundary, dev)
    plan.set_kernel(k)
    whole = plan.execute([ad], None, rd)
    plan.close()
    assert float((full - whole).abs().max() / whole.abs().max()) <= 5e-6
    a64, k64 = a.astype(np.float64), k.astype(np.float64)
    conv = orc.conv_reference_fast(a64, k64) if boundary == "reference" else orc.conv_same(a64, k64, fast=True)
    assert orc.rel_err_of_peak(full.cpu().numpy(), orc.density_correct(conv, rho, 1.0, 0.1, 0.0)) <= TOL


def test_eight_rank_slab_uses_the_180_point_transform():
    from pyvoxeldosimetry_b200.multi_gpu import slab_geometry

    for boundary in ("same", "reference"):
        g = slab_geometry((1024, 1024, 800), (51, 51, 51), boundary, 8, 3)
        assert g["n"][0] == 178 and g["ex"]["m"][0] == 180, g
    assert slab_geometry((1024, 1024, 800), (51, 51, 51), "same", 4, 1)["ex"]["m"][0] == 320
    assert slab_geometry((1024, 1024, 800), (51, 51, 51), "same", 2, 1)["ex"]["m"][0] == 576


def test_c5_full_size_eight_slabs_vs_whole_volume_and_properties():
    """BASELINE config 5 at full size (1024 x 1024 x 800, 51^3 Y90 kernel, zero boundary), 8 emulated ranks on one GPU:
    stitched slabs == whole-volume plan, sum conservation and the point response (size-independent properties), and -
    when the host has the memory for it - the float64 oracle of the whole volume."""
    from pyvoxeldosimetry_b200.engine import ConvPlan
    from pyvoxeldosimetry_b200.multi_gpu import SlabConvolver

    dev = torch.device("cuda:0")
    shape, ks, world = (1024, 1024, 800), (51, 51, 51), 8
    k = orc.y90_kernel(1.0, ks, "water").astype(np.float32)
    g = torch.Generator(device=dev).manual_seed(5)
    a = torch.zeros(shape, device=dev)
    rs = np.random.default_rng(5)
    for _ in range(32):  # boxes well inside the volume (the kernel reach is 25 voxels), one straddling every slab cut
        c = [int(rs.integers(70, n - 70)) for n in shape]
        h = [int(rs.integers(8, 40)) for _ in range(3)]
        a[c[0] - h[0] : c[0] + h[0], c[1] - h[1] : c[1] + h[1], c[2] - h[2] : c[2] + h[2]] += float(rs.uniform(1e5, 2e6))
    for cut in range(128, 1024, 128):
        a[cut - 6 : cut + 6, 500:530, 390:420] += 1.5e6
    p = (517, 300, 411)
    a[p] += 3e9  # a point source on top: its response is the kernel itself
    full = torch.empty(shape, device=dev)
    for rank in range(world):
        sc = SlabConvolver(shape, k, "same", device=dev, rank=rank, world=world)
        assert sc.plan.fft_shape == (180, 1152, 840)
        sc.fill_from_global(a)
        full[sc.lo : sc.hi] = sc(exchange=False)
        sc.check_device_errors()
        sc.plan.close()
        del sc
    torch.cuda.empty_cache()
    plan = ConvPlan(shape, ks, "same", dev)
    plan.set_kernel(k)
    whole = plan.execute([a])
    plan.check_device_errors()
    plan.close()
    peak = float(whole.abs().max())
    assert float((full - whole).abs().max()) / peak <= 5e-6
    # conservation: nothing leaves the volume (sources are > kernel reach from every face)
    want_sum = float(a.double().sum()) * float(k.astype(np.float64).sum())
    assert abs(float(full.double().sum()) - want_sum) / want_sum <= 1e-5
    # point response: subtracting the run without the point source leaves 3e9 * kernel around p
    a[p] -= 3e9
    plan = ConvPlan(shape, ks, "same", dev)
    plan.set_kernel(k)
    base = plan.execute([a])
    plan.close()
    resp = (whole - base)[p[0] - 25 : p[0] + 26, p[1] - 25 : p[1] + 26, p[2] - 25 : p[2] + 26].cpu().numpy() / 3e9
    assert np.max(np.abs(resp - k)) / k.max() <= 2e-5 * peak / 3e9 + 1e-4
    del base, plan
    if _mem_available_gb() >= 96.0:
        a[p] += 3e9
        ref = orc.conv_same(a.cpu().numpy().astype(np.float64), k.astype(np.float64), fast=True)
        assert orc.rel_err_of_peak(full.cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_two_process_halo_exchange_matches_the_whole_volume(transport):
    """The real exchange (one process per GPU, torchrun): mapped-peer-memory pulls and the NCCL send/recv fallback, both
    boundary modes, several epochs on one SlabConvolver.  Needs two GPUs (the driver's scaling boxes; skipped on one)."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SLAB_TRANSPORT=transport)
    port = 29531 + (transport == "nccl")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(repo, "scripts", "multi_gpu_check.py")], cwd=repo, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["transport"] == transport
    errs = [v for k, v in res.items() if k.startswith("slab_")]
    assert len(errs) == 3 and max(errs) < 1e-5


def test_stream_flags_copy_and_reserved_sms_on_one_gpu():
    """The building blocks of the peer transport on ONE device: a stream-ordered flag write releases a stream-ordered wait,
    pvd_copy_async copies device memory, and a plan with SMs reserved (smaller persistent grids, plain first launch) gives
    bit-identical results."""
    from pyvoxeldosimetry_b200._capi import PvdoseError
    from pyvoxeldosimetry_b200.engine import ConvPlan, get_lib

    dev = torch.device("cuda:0")
    lib = get_lib()
    flags = torch.zeros(4, dtype=torch.int32, device=dev)
    src = torch.arange(1000, dtype=torch.float32, device=dev)
    dst = torch.zeros_like(src)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize(dev)
    with torch.cuda.stream(s2):  # the waiter is enqueued FIRST and must hold the copy back until the flag is raised
        lib.stream_wait_flag_geq(flags[1:].data_ptr(), 7, s2.cuda_stream)
        lib.copy_async(dst.data_ptr(), src.data_ptr(), src.numel() * 4, s2.cuda_stream)
    assert not s2.query()
    with torch.cuda.stream(s1):
        lib.stream_write_flag(flags[1:].data_ptr(), 7, s1.cuda_stream)
    s2.synchronize()
    assert torch.equal(dst, src) and flags.tolist() == [0, 7, 0, 0]
    with pytest.raises(PvdoseError):
        lib.stream_write_flag(0, 1, 0)
    g = torch.Generator(device=dev).manual_seed(4)
    shape, ks = (180, 256, 256), (9, 9, 9)
    plan = ConvPlan(shape, ks, "reference", dev)
    plan.set_kernel(torch.rand(ks, device=dev, generator=g))
    a = torch.rand(shape, device=dev, generator=g)
    ref = plan.execute([a]).clone()
    lib.plan_reserve_sms(plan.handle, 32)
    got = plan.execute([a]).clone()
    lib.plan_reserve_sms(plan.handle, 0)
    assert torch.equal(ref, got)
    with pytest.raises(PvdoseError):
        lib.plan_reserve_sms(plan.handle, 10 ** 6)
    plan.check_device_errors()
    plan.close()
