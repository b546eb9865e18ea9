"""GPU tests of the code paths the default plan does not take at these sizes: the one-tile-per-CTA kernels
(PVD_ALGO_FFT_UNPIPELINED - the form the engine falls back to where the persistent TMA-pipelined tiles do not fit an SM)
and the any-length engine (PVD_FORCE_GENERIC=1, the library's only environment hook, read per plan).  Each must reproduce
the float64 oracle like the default path.  The cp.async staging that TMA replaced is reached without any switch: by
kernels whose rows are not 16-byte multiples (every pvd_plan_set_kernel), by the 256-point x pass and by cropped rows."""
import numpy as np
import pytest
import torch

from oracle import dose_oracle as orc

pytestmark = pytest.mark.gpu

CASES = (((512, 64, 400), (9, 7, 5), "reference"), ((192, 192, 256), (7, 7, 7), "reference"), ((160, 160, 224), (31, 31, 31), "same"),
         ((256, 40, 255), (5, 9, 6), "reference"))


@pytest.mark.parametrize("variant", ["default", "unpipelined", "generic"])
def test_alternative_paths_match_oracle(variant, monkeypatch):
    from pyvoxeldosimetry_b200 import _capi
    from pyvoxeldosimetry_b200.engine import ConvPlan

    monkeypatch.setenv("PVD_FORCE_GENERIC", "1" if variant == "generic" else "0")
    algo = _capi.ALGO_FFT_UNPIPELINED if variant == "unpipelined" else _capi.ALGO_FFT
    rng = np.random.default_rng(5)
    for shape, ks, boundary in CASES:
        a = rng.uniform(0, 1e3, shape).astype(np.float32)
        k = rng.uniform(0, 1.0, ks).astype(np.float32)
        rho = rng.choice([0.26, 1.04, 1.42], size=shape).astype(np.float32)
        plan = ConvPlan(shape, ks, boundary, "cuda:0", algo=algo)
        plan.set_kernel(k)
        out = plan.execute([torch.from_numpy(a).cuda()], None, torch.from_numpy(rho).cuda())
        plan.check_device_errors()
        a64, k64 = a.astype(np.float64), k.astype(np.float64)
        conv = orc.conv_reference_fast(a64, k64) if boundary == "reference" else orc.conv_same(a64, k64, fast=True)
        err = orc.rel_err_of_peak(out.cpu().numpy(), orc.density_correct(conv, rho, 1.0, 0.1, 0.0))
        plan.close()
        assert err <= 1e-4, (variant, shape, err)
