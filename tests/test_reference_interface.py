""""interface_wholebody_qref.py and the demo run unchanged" -- executed, not claimed.

tests/golden/interface_demo1.npz is the log of the reference's own demo_wholebody_qref.py + Interface, loaded from
/root/reference and run UNMODIFIED on the drop-in MPCWholeBody (tests/refshim/loader.py), with the CPU oracle doing the
solves.  CPU tier: (i) where the reference checkout exists the run is repeated and must reproduce the file; (ii) the
restated Interface (oracle/episode.py) and (iii) the drop-in class fed the recorded calls must agree with it.
GPU tier: the recorded calls replayed through the CUDA class, and the device-side episode against the real Interface's log."""
import os
import sys

import numpy as np
import pytest

from mobile_manipulator_mpc_b200 import _abi, scenarios
from oracle.episode import Episode

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
SHIM = os.path.join(HERE, "refshim")
FLAGS = ("move", "approach", "rotate", "move finish", "manipulate", "manipulate finish")


def _gold():
    return np.load(os.path.join(GOLD, "interface_demo1.npz"))


def _shim_import(name):
    sys.path.insert(0, SHIM)
    try:
        return __import__(name)
    finally:
        sys.path.remove(SHIM)


@pytest.mark.skipif(not os.path.isfile("/root/reference/interface_wholebody_qref.py"),
                    reason="the reference checkout only exists in the authoring container")
def test_reference_demo_and_interface_run_unchanged_on_the_dropin_class():
    sys.path.insert(0, GOLD)
    try:
        import make_interface_golden as mk
        o = mk.run()
    finally:
        sys.path.remove(GOLD)
        for m in ("make_interface_golden", "loader", "oracle_controller"):
            sys.modules.pop(m, None)
    g = _gold()
    assert o["x_log"].shape == (199, 9) and o["u_log"].shape == (198, 5)
    assert FLAGS[int(o["final_flag"])] == "manipulate finish"                  # the button is pushed
    for k in ("x_log", "u_log", "flags", "call_x_ref", "call_U", "call_flag", "call_Qd"):
        assert np.array_equal(o[k], g[k]), k
    # nothing of the reference or of the stand-ins leaks into the rest of the session
    assert "casadi" not in sys.modules and "interface_wholebody_qref" not in sys.modules


def test_restated_interface_matches_the_real_interface_log():
    """oracle/episode.py (the restatement the device-side state machine is tested against) vs. the real Interface."""
    g = _gold()
    x_start, tgt, planes = scenarios.demo_scenario(1)
    ep = Episode(float(g["dt"]), int(g["t_move"]), int(g["t_manipulate"]), x_start, tgt, scenarios.DEMO_CIRCLES, planes, N=int(g["N"]))
    xs, us, flags_before = [], [], []
    while ep.active and ep.steps < 400:
        flags_before.append(ep.flag)
        x_before = ep.state.copy()
        ep.step()
        xs.append(x_before)
        if ep.active:
            us.append(ep.u_latest[0].copy())
    assert ep.flag == "manipulate finish" and ep.steps == 199
    assert [FLAGS.index(f) for f in flags_before] == g["flags"].tolist()
    assert np.abs(np.array(xs) - g["x_log"]).max() < 1e-9
    assert np.abs(np.array(us) - g["u_log"]).max() < 1e-9
    assert np.abs(ep.local_pose_target - g["local_pose_target"]).max() < 1e-12
    assert np.abs(ep.traj_ref[-1, 6:] - g["q_target"]).max() < 1e-9
    # the weight switches and the one-time terminal equality, as the real Interface issued them
    assert g["call_flag"].tolist() == [0] * 42 + [1] * 156
    q = g["call_Qd"]
    assert (q[:66] == [25, 25, 0, 0, 0, 5, 5, 5, 5]).all() and (q[66:176] == [5, 5, 5, 0, 0, 1, 1, 1, 1]).all()
    assert (q[176:] == [500, 500, 500, 0, 0, 1, 1, 1, 1]).all()


def _replay(ctrl_cls, g, n=None, **kw):
    """Feed the calls the real Interface made to a controller class; returns U0 and cost per call."""
    from mobile_manipulator_mpc_b200.robot_models import MobileManipulator, Obstacles
    x_start, tgt, planes = scenarios.demo_scenario(1)
    pl = [(p[:3], p[3:].reshape(1, 3)) for p in planes]
    c = ctrl_cls(MobileManipulator(float(g["dt"])), [Obstacles(*o) for o in scenarios.DEMO_CIRCLES], pl, N=int(g["N"]), verbose=False, **kw)
    n = len(g["call_cost"]) if n is None else n
    u0, cost = [], []
    for i in range(n):
        if g["call_flag"][i] and not c.terminal_xy_eq:
            c.opti.subject_to(c.X[c.N, :2] == c.X_ref[c.N, :2])                    # interface_wholebody_qref.py:167
        if i == 0 or not np.array_equal(g["call_Qd"][i], g["call_Qd"][i - 1]):
            c.setWeight(Q=np.diag(g["call_Qd"][i]), P=np.diag(g["call_Pd"][i]))       # :175-177, :212-215
        x = g["call_x_init"][i].copy()
        assert c.u_latest is None or np.abs(c.u_latest - g["call_u_last"][i]).max() < 1e-3   # (state handling only: U* itself is compared below)
        if c.u_latest is not None:
            c.u_latest = g["call_u_last"][i].copy()                                   # replay: same U_last as recorded
        u0.append(c.solve(x, g["call_x_ref"][i], g["call_u_ref"][i]).copy())
        cost.append(c.cost)
    return np.array(u0), np.array(cost)


def test_dropin_class_replays_the_recorded_calls_cpu():
    g = _gold()
    oc = _shim_import("oracle_controller")
    u0, cost = _replay(oc.MPCWholeBody, g, n=60)
    assert np.abs(u0 - g["call_U"][:60, 0]).max() < 1e-12 and np.abs(cost - g["call_cost"][:60]).max() < 1e-9


@pytest.mark.gpu
def test_cuda_class_replays_the_calls_the_real_interface_made():
    """north_star tolerances on every one of the 198 solves of the reference's demo episode."""
    from mobile_manipulator_mpc_b200.controllers.mpc_wholebody_qref import MPCWholeBody
    g = _gold()
    u0, cost = _replay(MPCWholeBody, g)
    du = np.abs(u0 - g["call_U"][:, 0]).max(axis=1)
    rc = np.abs(cost - g["call_cost"]) / np.maximum(1e-12, np.abs(g["call_cost"]))
    assert (du < 1e-4).all(), (int((du >= 1e-4).sum()), np.argmax(du), du.max())
    assert (rc < 1e-5).all(), (int((rc >= 1e-5).sum()), np.argmax(rc), rc.max())


@pytest.mark.gpu
def test_device_episode_matches_the_real_interface_log():
    """csrc/mmpc_episode.cuh + the CUDA solver, free-running, against x_log / u_log / task flags of the real Interface."""
    from mobile_manipulator_mpc_b200.episodes import BatchedInterface
    g = _gold()
    x_start, tgt, planes = scenarios.demo_scenario(1)
    T = BatchedInterface(float(g["dt"]), int(g["t_move"]), int(g["t_manipulate"]), x_start[None], tgt[None], scenarios.DEMO_CIRCLES[None],
                         planes[None], N=int(g["N"]), mode=_abi.MODE_REFERENCE)
    xs, flags_before = [], []
    while T.steps < 400:
        flags_before.append(int(T.task.cpu()[0]))          # what the Interface prints at the top of timerCallback (:107)
        xs.append(T.x.cpu().numpy()[0].copy())             # x_log.append(current_state) (:109)
        if T.step() == 0:
            break
    n = len(g["x_log"])
    assert len(xs) == n, (len(xs), n)
    assert int(T.task.cpu()[0]) == _abi.TASK_FINISHED
    assert [FLAGS.index(_abi.TASK_NAMES[f]) for f in flags_before] == g["flags"].tolist()
    assert np.abs(np.array(xs) - g["x_log"]).max() < 1e-5
    T.close()
