"""the base-only sibling controller MPCBase (reference: controllers/mpc_base.py) behind the same C ABI
(MmpcConfig.model = MMPC_MODEL_BASE).  oracle/nlp_base.py is pinned to the reference's own reset() in tests/test_reference_rows.py;
here: the oracle solver against independent SciPy solutions of that restatement, the kernel sources against the oracle (CPU
emulation), the drop-in class surface, and -- GPU tier -- the CUDA path against the oracle."""
import os
import sys

import numpy as np
import pytest

from mobile_manipulator_mpc_b200 import _abi, scenarios
from mobile_manipulator_mpc_b200.robot_models import Base, Obstacles
from oracle.nlp_base import NLPBase

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, os.path.join(HERE, "refshim"))
import oracle_controller as OC  # noqa: E402
sys.path.remove(os.path.join(HERE, "refshim"))


def _ctrl(cls, b, **kw):
    return cls(Base(b["dt"]), [Obstacles(*c) for c in b["circles"][0]], N=b["N"], **kw)


def _check_against_oracle(out, ref, b, max_unconverged=0, max_outliers=0):
    both = (out["status"] == 0) & (ref["status"] == 0)
    assert int((out["status"] != 0).sum()) <= max_unconverged and int((ref["status"] != 0).sum()) <= max_unconverged, (out["status"], ref["status"])
    rel = np.abs(out["cost"] - ref["cost"]) / np.maximum(np.abs(ref["cost"]), 1e-300)
    du0 = np.abs(out["U"][:, 0] - ref["U"][:, 0]).max(axis=1)
    outliers = np.nonzero(both & ((rel > 1e-5) | (du0 > 1e-4)))[0]
    assert len(outliers) <= max_outliers, (outliers, rel[outliers], du0[outliers])
    for i in np.nonzero(both)[0][:32]:
        P = NLPBase(b["N"], b["dt"], b["x_init"][i], b["x_ref"][i], b["u_ref"][i], b["circles"][i])
        assert P.violation(P.pack(out["X"][i], out["U"][i], out["s"][i])) <= 1e-6, i
        assert abs(P.cost(P.pack(out["X"][i], out["U"][i], out["s"][i])) - out["cost"][i]) <= 1e-9 * abs(out["cost"][i])
    return both


def test_class_surface_and_warm_start():
    b = scenarios.make_base_batch(1)
    c = _ctrl(OC.MPCBase, b)
    assert c.N == 10 and c.dt == 0.1 and c.base_radius() == 0.4 and c.ulim.shape == (2, 2) and c.xlim.shape == (2, 5)
    assert c.X_guess is None and c.U_guess is None
    assert c.weights["Qd"].tolist() == [5, 5, 0, 0, 0, 1, 0, 0, 0] and c.weights["Rd"].tolist() == [1, 1, 1, 1, 1] and c.weights["S"] == 1e5
    u = c.solve(b["x_init"][0].copy(), b["x_ref"][0], b["u_ref"][0])
    assert u.shape == (2,) and u.dtype == np.float64 and c.X_guess.shape == (11, 6) and c.U_guess.shape == (10, 2)
    it1 = int(c.last_info["iters"][0])
    x1 = np.asarray(c.f_dynamics(b["x_init"][0], u)).reshape(6)
    c.solve(x1, b["x_ref"][0], b["u_ref"][0])                       # second step: X and U start from the previous solution (:196-201)
    assert int(c.last_info["iters"][0]) <= it1
    assert abs(c.angleDiff(3.1, -3.1) - (6.2 - 2 * np.pi)) < 1e-15
    c.setWeight(Q=np.diag([1, 1, 2.0, 0, 0, 1]))
    assert c.weights["Qd"][2] == 2.0
    with pytest.raises(_abi_error()):
        from mobile_manipulator_mpc_b200.controllers.mpc_base import MPCBase
        _ctrl(MPCBase, b).solve(b["x_init"][0].copy(), b["x_ref"][0], b["u_ref"][0])     # no CUDA library / device here: no CPU fallback


def _abi_error():
    from mobile_manipulator_mpc_b200._lib import MmpcError
    return (MmpcError, OSError)


def test_oracle_solver_matches_scipy_on_the_restated_nlp():
    g = np.load(os.path.join(GOLD, "base_N10_slsqp.npz"))
    b = dict(N=int(g["N"]), dt=float(g["dt"]), x_init=g["x_init"], x_ref=g["x_ref"], u_ref=g["u_ref"], circles=g["circles"])
    c = _ctrl(OC.MPCBase, b)
    out = c.solve_batch(b["x_init"], b["x_ref"], b["u_ref"], circles=b["circles"])
    assert (out["status"] == 0).all()
    for i in (0, 1, 3, 5):     # the instances SLSQP finished on (2 and 4 start inside a circle; SLSQP stops with status 8 there)
        assert abs(out["cost"][i] - g["cost"][i]) <= 1e-6 * g["cost"][i], i
        assert np.abs(out["U"][i, 0] - g["u0"][i]).max() < 1e-4, i
    for i in range(6):
        P = NLPBase(b["N"], b["dt"], b["x_init"][i], b["x_ref"][i], b["u_ref"][i], b["circles"][i])
        assert P.violation(P.pack(out["X"][i], out["U"][i], out["s"][i])) <= 1e-6   # (SLSQP's points for 2 and 4 violate by 5e-4 / 0.15)


@pytest.mark.parametrize("n_obs", [3, 16])
def test_kernel_sources_match_oracle_emulated(n_obs):
    from tests.emu import emu
    b = scenarios.make_base_batch(6, n_obs=n_obs, seed=8)
    ref = _ctrl(OC.MPCBase, b).solve_batch(b["x_init"], b["x_ref"], b["u_ref"], circles=b["circles"])

    class Emu(OC.MPCBase):
        BACKEND = staticmethod(lambda bb, cfg: emu.solve(bb, cfg, kernel="staged"))
    out = _ctrl(Emu, b).solve_batch(b["x_init"], b["x_ref"], b["u_ref"], circles=b["circles"])
    _check_against_oracle(out, ref, b)
    assert (np.abs(out["iters"] - ref["iters"]) <= 2).all()
    assert np.abs(out["X"][:, :, :] - ref["X"]).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["staged", "resident"])
@pytest.mark.parametrize("n_obs,B", [(3, 256), (16, 256)])
def test_gpu_mpc_base_matches_oracle(n_obs, B, kernel):
    from mobile_manipulator_mpc_b200.controllers.mpc_base import MPCBase
    b = scenarios.make_base_batch(B, n_obs=n_obs, seed=9)
    ref = _ctrl(OC.MPCBase, b).solve_batch(b["x_init"], b["x_ref"], b["u_ref"], circles=b["circles"])
    c = _ctrl(MPCBase, b, batch=B)
    c.solve_batch(b["x_init"][:1], b["x_ref"][:1], b["u_ref"][:1], circles=b["circles"][:1])   # creates the solver
    c._solver.set_kernel(kernel)
    out = c.solve_batch(b["x_init"], b["x_ref"], b["u_ref"], circles=b["circles"])
    both = _check_against_oracle(out, ref, b, max_unconverged=1, max_outliers=1)
    assert both.sum() >= B - 1
    # the single-instance call and the warm start
    u = c.solve(b["x_init"][0].copy(), b["x_ref"][0], b["u_ref"][0])
    assert np.abs(u - ref["U"][0, 0]).max() < 1e-4 and c.X_guess.shape == (b["N"] + 1, 6)
