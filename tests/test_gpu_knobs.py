"""GPU tests of the alternative code paths behind the environment knobs: the cp.async staging
that TMA replaced, normal launches instead of programmatic dependent launch, the one-tile-per-CTA x pass and the
dual-group x pass.  Each combination runs in its own process (the knobs are read once) on a grid whose three lengths
are on the size-specialised menu, and must reproduce the float64 oracle like the default path."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
sys.path.insert(0, %r)
import numpy as np, torch
from oracle import dose_oracle as orc
from pyvoxeldosimetry_b200.engine import ConvPlan
rng = np.random.default_rng(5)
worst = 0.0
for shape, ks, boundary in (((512, 64, 400), (9, 7, 5), 'reference'), ((192, 192, 256), (7, 7, 7), 'reference'), ((160, 160, 224), (31, 31, 31), 'same')):
    a = rng.uniform(0, 1e3, shape).astype(np.float32)
    k = rng.uniform(0, 1.0, ks).astype(np.float32)
    rho = rng.choice([0.26, 1.04, 1.42], size=shape).astype(np.float32)
    plan = ConvPlan(shape, ks, boundary, 'cuda:0')
    plan.set_kernel(k)
    out = plan.execute([torch.from_numpy(a).cuda()], None, torch.from_numpy(rho).cuda())
    plan.check_device_errors()
    conv = orc.conv_reference_fast(a.astype(np.float64), k.astype(np.float64)) if boundary == 'reference' else orc.conv_same(a.astype(np.float64), k.astype(np.float64), fast=True)
    worst = max(worst, orc.rel_err_of_peak(out.cpu().numpy(), orc.density_correct(conv, rho, 1.0, 0.1, 0.0)))
    plan.close()
print('WORST', worst)
assert worst <= 1e-4, worst
""" % REPO

KNOBS = [
    {},
    {"PVD_TMA": "0", "PVD_TMA_ROWS": "0"},
    {"PVD_TMA_DEN": "0", "PVD_PDL": "0"},
    {"PVD_P3_LOOP": "0"},
    {"PVD_P3_DUAL": "1"},
    {"PVD_NO_PIPE": "1"},
]


@pytest.mark.parametrize("knobs", KNOBS, ids=lambda k: ",".join(f"{a}={b}" for a, b in k.items()) or "default")
def test_alternative_paths_match_oracle(knobs):
    env = dict(os.environ, **knobs)
    res = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "WORST" in res.stdout
