"""CPU tier: the CUDA solver *sources* (csrc/mmpc_staged.cuh + mmpc_team.cuh, mmpc_lane.cuh,
mmpc_solver.cuh) executed by the CPU emulators of tests/emu against the independent dense oracle
(oracle/mmpc_oracle.c).  This is how the kernels are debugged on the GPU-less authoring box; the
real parity tests are the -m gpu ones."""
import numpy as np
import pytest

from mobile_manipulator_mpc_b200 import _abi, scenarios
from oracle import solver
from tests.emu import emu


def _compare(batch, kernel, tol_cost=1e-6, tol_u=1e-4):
    cfg = solver.config_from_batch(batch, mode=_abi.MODE_CLEAN)
    o = solver.solve(batch, cfg=cfg, threads=4)
    e = emu.solve(batch, cfg, kernel=kernel)
    both = (o["status"] == 0) & (e["status"] == 0)
    assert both.mean() >= 0.9
    rel = np.abs(o["cost"] - e["cost"])[both] / np.abs(o["cost"][both])
    du = np.abs(o["U"][:, 0] - e["U"][:, 0]).max(axis=1)[both]
    assert (rel < tol_cost).mean() >= 0.95 and (du < tol_u).mean() >= 0.95
    return o, e


# staged = phase bodies + 16-lane team Riccati (the product default); staged_thread = phase bodies +
# one-thread-per-instance Riccati; lane / warp = the two persistent single-kernel solvers
@pytest.mark.parametrize("kernel", ["staged", "staged_thread", "lane", "warp"])
def test_config1_bit_level_agreement(kernel):
    b = scenarios.make_batch(1, 1)
    o, e = _compare(b, kernel)
    assert o["iters"][0] == e["iters"][0]
    assert np.abs(o["X"] - e["X"]).max() < 1e-9 and np.abs(o["U"] - e["U"]).max() < 1e-9


@pytest.mark.parametrize("kernel,B3,B5", [("staged", 4, 2), ("staged_thread", 12, 3), ("lane", 12, 3), ("warp", 12, 3)])
def test_config3_and_moving_obstacles(kernel, B3, B5):
    _compare(scenarios.make_batch(3, B3), kernel)
    _compare(scenarios.make_batch(5, B5), kernel)


def test_staged_rounds_and_list_compaction():
    """The staged solver finishes in max(iterations + line-search retries) + 1 rounds, every instance
    leaves the lists exactly once, and the thread and team Riccati agree to rounding."""
    b = scenarios.make_batch(2, 6)
    cfg = solver.config_from_batch(b, mode=_abi.MODE_CLEAN)
    a = emu.solve(b, cfg, kernel="staged_thread")
    t = emu.solve(b, cfg, kernel="staged")
    assert (a["status"] == 0).all() and (t["status"] == 0).all()
    assert a["rounds"] >= a["iters"].max() + 1 and a["rounds"] <= a["iters"].max() + 60
    assert (a["iters"] == t["iters"]).all()
    assert np.abs(a["U"] - t["U"]).max() < 1e-8 and np.abs(a["cost"] - t["cost"]).max() < 1e-8 * np.abs(a["cost"]).max()


@pytest.mark.parametrize("kernel", ["staged", "staged_thread", "staged_fat"])
def test_terminal_xy_equality_flag(kernel):
    """flags bit 0 = opti.subject_to(X[N,:2] == X_ref[N,:2]) (interface_wholebody_qref.py:167)."""
    b = scenarios.approach_instance()
    cfg = solver.config_from_batch(b, mode=_abi.MODE_CLEAN)
    o = solver.solve(b, cfg=cfg, threads=1)
    e = emu.solve(b, cfg, kernel=kernel)
    assert o["status"][0] == 0 and e["status"][0] == 0
    assert np.abs(e["X"][0, -1, :2] - b["x_ref"][0, -1, :2]).max() < 1e-8          # the equality holds
    assert abs(e["cost"][0] - o["cost"][0]) <= 1e-9 * abs(o["cost"][0])
    assert np.abs(e["U"][0, 0] - o["U"][0, 0]).max() < 1e-7
    free = dict(b); free["flags"] = None                                            # and it binds: without it x_N differs
    f = emu.solve(free, cfg, kernel=kernel)
    assert np.abs(f["X"][0, -1, :2] - b["x_ref"][0, -1, :2]).max() > 1e-3
