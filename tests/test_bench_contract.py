"""Host-side checks of bench.py's contract pieces that need no GPU: both arms of a workload print the SAME `config`
object, roofline.traffic is taken only from an ncu capture stamped with the running library's build id, and the
`--impl reference` arm (the reference's own class from oracle/_ref, or the oracle port) prints a complete line."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def test_config_is_identical_for_both_arms_and_names_the_workload():
    import bench

    for name in ("c1", "c2", "c3"):
        wl = bench.WORKLOADS[name]
        a, b = bench.arm_config(wl, "reference"), bench.arm_config(wl, "reference")
        assert a == b and a["workload"] == wl["desc"] and "model" not in a
        assert ("exceed" in a["l2"]) == (name != "c1")  # C1 (48^3) fits the L2 and says so


def test_dram_traffic_is_reported_only_for_the_running_build():
    import bench
    import __graft_entry__ as g

    p = os.path.join(REPO, "profiles", "r02_dram_traffic_c3.json")
    stamp = json.load(open(p))["build_id"]
    traffic, src = bench.dram_traffic_for(stamp, "c3", "reference")
    assert traffic and traffic > 1.2e9 and src == "r02_dram_traffic_c3.json"  # >= the algorithmic 12 B/voxel
    traffic, src = bench.dram_traffic_for("0123456789abcdef", "c3", "reference")
    assert traffic is None and src.startswith("stale")
    assert bench.dram_traffic_for(stamp, "c1", "reference") == (None, "no capture for this workload")
    # the committed captures should belong to the committed kernels; when the CUDA sources have moved on, bench.py reports
    # traffic = null ("stale") until scripts/final_gpu.sh is re-run - flagged here, not failed (no GPU on this side)
    if stamp != g.source_hash():
        import warnings

        warnings.warn("CUDA sources changed after the ncu traffic capture: re-run scripts/final_gpu.sh")
        assert bench.dram_traffic_for(g.source_hash(), "c3", "reference")[0] is None
    for f in ("r02_dram_traffic_c3_same.json", "r02_dram_traffic_c2_reference.json"):
        assert json.load(open(os.path.join(REPO, "profiles", f)))["build_id"] == stamp


def test_reference_arm_prints_a_complete_line_on_the_cpu():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, cwd=REPO, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "dose_volumes_per_sec" and line["unit"] == "volumes/s"
    assert line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench

    assert line["config"] == bench.arm_config(bench.WORKLOADS["c1"], "reference")
