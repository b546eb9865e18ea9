"""The reference's own example scripts, UNMODIFIED, against the drop-in ("run them with a stub
matplotlib, not edit them").  The scripts are executed from oracle/_ref/examples - byte-identical copies of
/root/reference/examples made by oracle/ref_loader.py (sha256 manifest; git-ignored, shipped to the GPU box like a built
.so) - or straight from /root/reference/examples when that exists.  Each runs in its own process with `pyvoxeldosimetry`
resolving to this repository's alias package and matplotlib replaced by a permissive stub (absent in this image); the
dose arrays the script computed are then compared with the float64 oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import dose_oracle as orc
from oracle import ref_loader

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _example(name):
    for d in (os.path.join(ref_loader.REF_DIR, "examples"), "/root/reference/examples"):
        p = os.path.join(d, name)
        if os.path.exists(p):
            return p
    return None


RUNNER = r"""
import runpy, sys, os
import numpy as np
sys.path.insert(0, {repo!r})
from oracle.ref_loader import Stub            # test infrastructure: the permissive matplotlib stand-in
for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[m] = Stub(m)
import pyvoxeldosimetry
assert "pyvoxeldosimetry_b200" in sys.modules and os.path.dirname(pyvoxeldosimetry.__file__).startswith({repo!r}), pyvoxeldosimetry.__file__
g = runpy.run_path({script!r}, run_name="__main__")
out = {{}}
for k, v in g.items():
    if isinstance(v, np.ndarray) and v.ndim == 3:
        out[k] = v
res = g.get("result")
if res is not None:
    if res.absorbed_dose is not None:
        out["result_absorbed_dose"] = np.asarray(res.absorbed_dose)
    for i, r in enumerate(res.dose_rate_maps):
        out["result_rate_%d" % i] = np.asarray(r)
if "activity_maps" in g:
    for i, a in enumerate(g["activity_maps"]):
        out["activity_maps_%d" % i] = np.asarray(a)
np.savez({npz!r}, **out)
"""


def _run(script, tmp_path):
    npz = str(tmp_path / "vars.npz")
    code = RUNNER.format(repo=REPO, script=script, npz=npz)
    res = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    return np.load(npz), res.stdout


def test_vendored_examples_are_byte_identical_to_the_reference():
    if not ref_loader.available() or not os.path.isdir("/root/reference/examples"):
        pytest.skip("needs both oracle/_ref and /root/reference (build container only)")
    assert ref_loader.verify_against_source() == []


@pytest.mark.gpu
def test_single_timepoint_y90_physical_decay_example(tmp_path):
    script = _example("single_timepoint_y90_physical_decay.py")
    if script is None:
        pytest.skip("reference examples not available (oracle/_ref not built)")
    v, out = _run(script, tmp_path)
    a = v["activity_map"]
    assert a.shape == (48, 48, 48) and a.dtype == np.float64 and "Total activity: 4.22e+09 Bq" in out
    k = orc.y90_kernel(1.0, (64, 64, 64), "water")                       # the calculator's hard-coded 64^3 grid
    rate = orc.conv_reference(a, k.astype(np.float32).astype(np.float64))
    # the example reads result.absorbed_dose for ONE time point (None in the reference): physical-decay integral of the rate
    want = rate * (64.1 * 3600.0 / np.log(2.0))
    assert orc.rel_err_of_peak(v["dose_rate"], want) <= 1e-4
    assert orc.rel_err_of_peak(v["result_rate_0"], rate) <= 1e-4
    assert tuple(np.unravel_index(np.argmax(v["result_rate_0"]), a.shape)) == (8, 8, 8)  # (24 + 32) mod 48: SURVEY 0.3
    assert "Maximum dose rate:" in out and "nan" not in out.lower()


@pytest.mark.gpu
def test_kernel_convolution_example(tmp_path):
    script = _example("kernel_convolution_example.py")
    if script is None:
        pytest.skip("reference examples not available (oracle/_ref not built)")
    v, out = _run(script, tmp_path)
    a = v["activity_map"]
    assert a.shape == (64, 64, 64) and "Total activity:" in out
    from pyvoxeldosimetry_b200.data.dose_kernels.kernel_factory import KernelFactory

    k = np.asarray(KernelFactory().get_kernel("F18", "water", 1.0, (64, 64, 64)), dtype=np.float64)  # F18 has no runnable reference generator
    want = orc.conv_reference(a, k.astype(np.float32).astype(np.float64))
    assert orc.rel_err_of_peak(v["dose_rate"], want) <= 1e-4
    assert np.isfinite(v["dose_rate"]).all()


@pytest.mark.gpu
def test_time_integrated_dose_example(tmp_path):
    script = _example("time_integrated_dose.py")
    if script is None:
        pytest.skip("reference examples not available (oracle/_ref not built)")
    v, out = _run(script, tmp_path)
    tp = [0, 24, 48, 72, 96, 120]
    maps = [v["activity_maps_%d" % i] for i in range(len(tp))]
    k = orc.lu177_kernel(1.0, (64, 64, 64), "water").astype(np.float32).astype(np.float64)
    # default integration_mode 'activity': trapezoid of the activity in the caller's unit (hours), then ONE convolution
    acc = orc.integrate_activity_trapezoid(maps, tp)
    want = orc.conv_reference(acc, k)
    assert orc.rel_err_of_peak(v["result_absorbed_dose"], want) <= 1e-4
    for i in range(len(tp)):  # the example indexes result.dose_rate_maps[i] ([] in the reference)
        assert orc.rel_err_of_peak(v["result_rate_%d" % i], orc.conv_reference(maps[i], k)) <= 1e-4
    assert "Maximum absorbed dose:" in out and "mode: multi_timepoint_activity" in out
    assert os.path.isdir(tmp_path / "time_integrated_results")
