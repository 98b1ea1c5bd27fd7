"""GPU parity of the remaining EOSApply variants: EOS_PR and EOS_THERMO
(lbm_eos.F90:231-349) against the oracle, fused and split kernels, and the reference's
"PR EOS inner sqrt went negative" stop."""
import numpy as np
import pytest

import cases
import gpu_util
from taxila_lbm_b200 import capi
from taxila_lbm_b200 import config as tc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("split", [False, True])
def test_peng_robinson_and_thermo_eos(monkeypatch, split):
    if split:
        monkeypatch.setenv("TXG_SPLIT", "1")
    cfg, walls, rho = cases.eos_pr_thermo_3d(32)
    steps = 30
    o = cases.run_oracle(cfg, walls, rho, steps)
    assert o.eos_bad() == 0
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(steps)
    flow.synchronize()
    fi, r, u, F = gpu_util.fields(flow)
    fluid = walls == 0
    errs = {"fi": gpu_util.rel_err(fi, o.fi()), "rho": gpu_util.rel_err(r[fluid], o.rho()[fluid]),
            "u": gpu_util.rel_err(u[fluid], o.u()[fluid]), "forces": gpu_util.rel_err(F[fluid], o.forces()[fluid])}
    assert all(v <= 1e-10 for v in errs.values()), errs
    rhot, prs, velt = flow.update_diagnostics()
    ort, opr, ovt = o.diagnostics()
    assert gpu_util.rel_err(prs[fluid], opr[fluid]) <= 1e-10
    assert gpu_util.rel_err(velt[fluid], ovt[fluid]) <= 1e-10
    flow.close()


def test_peng_robinson_negative_root_is_an_error():
    """A repulsive self-interaction makes 2 (p_EOS - rho/3) / (c_0 g_mm) negative: LBMError in the reference
    (lbm_eos.F90:337-341, ierr = 58), an error code at the next synchronising call here."""
    cfg, walls, rho = cases.eos_pr_thermo_3d(16)
    cfg.gf[0][0] = +0.5
    with pytest.raises(capi.TaxilaGpuError) as e:
        flow = gpu_util.make_flow(cfg, walls, rho)
        flow.synchronize()
    assert e.value.code == 58 and "PR EOS" in str(e.value)
