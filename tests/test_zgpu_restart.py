"""Restart / IC-from-file on the host mirror; B200: profiles/r1s_restart_test.log."""
import numpy as np
import pytest

import cases
import gpu_util
import taxila_lbm_b200 as tx
from taxila_lbm_b200 import geometry as geo

pytestmark = pytest.mark.gpu


def test_restart_from_output_file_continues_bit_for_bit(tmp_path):
    """-restart (LBMInitializeStateRestarted, lbm.F90:524-544): a run restarted from the fi file that
    FlowOutputDiagnostics wrote continues exactly like the run that wrote it."""
    cfg, walls, rho = cases.porous_3d(24, rmin=3.0, rmax=6.0)
    a = gpu_util.make_flow(cfg, walls, rho)
    a.step(15)
    prefix = str(tmp_path / "run_")
    a.output_diagnostics(prefix, 3, rho=False, velt=False, rhot=False, prs=False)
    a.step(10)
    b = tx.Flow(cfg)
    b.walls_set_values(geo.ghosted(walls, cfg.stencil_size_rho, cfg.periodic, 3, wall_ghost=True))
    b.initialize_state_restarted(prefix, 3)
    b.update_moments()
    b.step(10)
    assert np.array_equal(a.get_fi(), b.get_fi())
    a.close()
    b.close()
