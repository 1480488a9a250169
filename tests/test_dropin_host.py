"""CPU tier: host-side mirror of the reference interface (no GPU): constructor surface, attributes
the Interface touches, opti shim, angleDiff, robot models, scenario generators."""
import numpy as np
import pytest

from mobile_manipulator_mpc_b200 import scenarios
from mobile_manipulator_mpc_b200.controllers.mpc_wholebody_qref import MPCWholeBody
from mobile_manipulator_mpc_b200.robot_models import MobileManipulator, Obstacles
from oracle import model as M


def _controller(N=20):
    robot = MobileManipulator(0.1)
    _, _, planes = scenarios.demo_scenario(1)
    obst = [Obstacles(*c) for c in scenarios.DEMO_CIRCLES]
    pl = [(p[:3], p[3:].reshape(1, 3)) for p in planes]
    return MPCWholeBody(robot, obst, pl, N=N, verbose=False)


def test_constructor_and_attribute_surface():
    c = _controller()
    assert c.N == 20 and c.dt == 0.1 and c.base_radius == 0.4
    assert c.x_guess is None and c.u_latest is None
    assert c.ulim.shape == (2, 5) and c.xlim.shape == (2, 9) and c.dulim.shape == (2, 5)
    assert [o.radius for o in c.obstacle_list] == [0.6, 0.6, 0.1]
    assert c.weights["Qd"].tolist() == [25, 25, 0, 0, 0, 5, 5, 5, 5] and c.weights["S"] == 1e5
    c.setWeight(P=np.diag([5, 5, 5, 0, 0, 1, 1, 1, 1]), Q=np.diag([5, 5, 5, 0, 0, 1, 1, 1, 1]))  # interface :175-177
    assert c.weights["Qd"].tolist() == [5, 5, 5, 0, 0, 1, 1, 1, 1]
    with pytest.raises(NotImplementedError):
        c.setWeight(Q=np.ones((9, 9)))


def test_opti_shim_recognises_the_interface_mutation():
    c = _controller()
    N = c.N
    c.opti.subject_to(c.X[N, :2] == c.X_ref[N, :2])        # interface_wholebody_qref.py:167
    assert c.terminal_xy_eq is True
    with pytest.raises(NotImplementedError):
        c.opti.subject_to(c.X[0, :2] == c.X_ref[N, :2])
    c.reset()
    assert c.terminal_xy_eq is False


def test_angle_diff_matches_oracle():
    c = _controller()
    rng = np.random.default_rng(0)
    for a, b in rng.uniform(-7, 7, (200, 2)):
        assert abs(c.angleDiff(a, b) - M.angle_diff(a, b)) < 1e-15
    assert abs(c.angleDiff(1.0, 2.0)) == 1.0   # abs()-able float, interface :194


def test_robot_models_match_oracle():
    robot = MobileManipulator(0.1)
    rng = np.random.default_rng(1)
    for _ in range(20):
        x = rng.uniform(-2, 2, 9); u = rng.uniform(-1, 1, 5)
        pe, j2, j3 = robot.forward_tranformation(x)
        ope, oj2, oj3 = M.forward_transformation(x)
        assert pe.shape == (1, 4) and j2.shape == (1, 3)
        assert np.allclose(pe[0], [float(v) for v in ope], atol=1e-15)
        assert np.allclose(j2[0], oj2, atol=1e-15) and np.allclose(j3[0], oj3, atol=1e-15)
        xc = x.copy()
        nxt = np.asarray(robot.f_kinematics(xc, u)).squeeze()
        assert np.allclose(nxt, M.f_kinematics(x, u, 0.1), atol=1e-15)
        assert np.allclose(xc[6:], nxt[6:])     # the reference's in-place q += q_dot*dt is reproduced
    e = robot.manipulator.forward_tranformation(np.zeros(3))[0]
    assert np.allclose(e, [[0.088, 0, 0.593]])


def test_inverse_transformation_known_answer():
    robot = MobileManipulator(0.1)
    q = robot.manipulator.inverse_transformation(np.array([-np.pi / 4, -np.pi, np.pi]), np.array([0.6, 0.0, 0.1]))
    e = robot.manipulator.forward_tranformation(q)[0][0]
    assert abs(e[0] - 0.6) < 1e-6 and abs(e[2] - 0.1) < 1e-6
    assert -np.pi / 2 <= q[0] <= np.pi / 2 and -3 * np.pi / 4 <= q[1] <= 0 and 0 <= q[2] <= 1.5 * np.pi


def test_obs_avoid_matches_oracle():
    c = _controller()
    x = np.array([2.0, 1.5, 0, 0, 0, 0, 0, 0, 0.0])
    assert np.allclose(c.obsAvoid(c.obstacle_list, x), M.circle_rows(x, scenarios.DEMO_CIRCLES), atol=1e-15)


def test_scenario_generators():
    b = scenarios.make_batch(1, 1)
    assert b["x_ref"].shape == (1, 21, 9) and b["u_ref"].shape == (1, 20, 5)
    assert np.allclose(b["x_ref"][0, -1, :2], [2.0, 2.0])        # row 20 of the 51-point linspace to (5,5)
    b2 = scenarios.make_batch(2, 64)
    assert b2["x_init"].shape == (64, 9) and (b2["x_init"][:, 7] <= 0).all() and (b2["x_init"][:, 8] >= 0).all()
    b3 = scenarios.make_batch(3, 64)
    assert b3["circles"].shape == (64, 16, 3) and set(b3["n_pl_inst"].tolist()) == {2, 3}
    d = np.linalg.norm(b3["circles"][:, :, :2] - b3["x_init"][:, None, :2], axis=2)
    assert (d >= b3["circles"][:, :, 2] + 0.5).all()
    b5 = scenarios.make_batch(5, 8)
    assert b5["circles"].shape == (8, 41, 16, 3) and b5["obs_per_stage"] == 1
    assert np.array_equal(scenarios.make_batch(3, 16)["x_init"], scenarios.make_batch(3, 16)["x_init"])  # seeded


def test_local_window_semantics():
    x_start, tgt, _ = scenarios.demo_scenario(1)
    ref, uref = scenarios.global_plan_2d(x_start, scenarios.base_target(x_start, tgt), 5, 0.1)
    assert ref.shape == (51, 9) and uref.shape == (50, 5)
    st = ref[40].copy()
    xr, ur = scenarios.local_window(ref, uref, st, [0, 1], 20)
    assert np.allclose(xr[0], ref[40]) and np.allclose(xr[10], ref[50]) and np.allclose(xr[20], ref[50])  # repeats last row
    assert xr.shape == (21, 9) and ur.shape == (20, 5)
