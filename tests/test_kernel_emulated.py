"""CPU tier: the CUDA solver kernel *source* (csrc/mmpc_solver.cuh) executed by the 32-lane CPU
emulator (tests/emu) against the independent dense oracle (oracle/mmpc_oracle.c).  This is how the
kernel is debugged on the GPU-less authoring box; the real parity tests are the -m gpu ones."""
import numpy as np

from mobile_manipulator_mpc_b200 import _abi, scenarios
from oracle import solver
from tests.emu import emu


def _compare(batch, tol_cost=1e-6, tol_u=1e-4):
    cfg = solver.config_from_batch(batch, mode=_abi.MODE_CLEAN)
    o = solver.solve(batch, cfg=cfg, threads=4)
    e = emu.solve(batch, cfg)
    both = (o["status"] == 0) & (e["status"] == 0)
    assert both.mean() >= 0.9
    rel = np.abs(o["cost"] - e["cost"])[both] / np.abs(o["cost"][both])
    du = np.abs(o["U"][:, 0] - e["U"][:, 0]).max(axis=1)[both]
    assert (rel < tol_cost).mean() >= 0.95 and (du < tol_u).mean() >= 0.95
    return o, e


def test_config1_bit_level_agreement():
    b = scenarios.make_batch(1, 1)
    o, e = _compare(b)
    assert o["iters"][0] == e["iters"][0]
    assert np.abs(o["X"] - e["X"]).max() < 1e-9 and np.abs(o["U"] - e["U"]).max() < 1e-9


def test_config3_and_moving_obstacles():
    _compare(scenarios.make_batch(3, 12))
    _compare(scenarios.make_batch(5, 3))
