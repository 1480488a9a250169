"""CPU tier: the expression-level oracle (oracle/model.py) against the reference's own artefacts:
the sympy DH derivation (tests/golden/fk_dh_golden.npz, made by make_fk_golden.py from
/root/reference/utils/dh_to_kinematics.py) and the known answers of SURVEY.md 8(a)/(c)."""
import os

import numpy as np

from oracle import model as M

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fk_matches_reference_dh_derivation():
    g = np.load(os.path.join(GOLD, "fk_dh_golden.npz"))
    (ex, ez), (j2x, j2z), (j3x, j3z) = M.arm_fk(g["q"])
    for got, ref in (((ex, ez), g["endpoint"]), ((j2x, j2z), g["joint2"]), ((j3x, j3z), g["joint3"])):
        assert np.allclose(got[0], ref[:, 0], rtol=1e-12, atol=1e-14)
        assert np.allclose(ref[:, 1], 0.0, atol=1e-15)          # planar arm: y == 0
        assert np.allclose(got[1], ref[:, 2], rtol=1e-12, atol=1e-14)


def test_fk_known_answers():
    (ex, ez), (j2x, j2z), (j3x, j3z) = M.arm_fk(np.zeros(3))
    assert np.allclose([ex, ez, j2x, j2z, j3x, j3z], [0.088, 0.593, 0.0825, 0.316, 0.0, 0.7], atol=1e-15)
    x = np.array([0, 0, 0, 0, 0, 0, -np.pi / 4, -np.pi, np.pi])       # demo_wholebody_qref.py:19
    pe, j2, j3 = M.forward_transformation(x)
    assert np.allclose(pe[:3], [0.29564170234784237, 0.0, 0.9941543289325507], atol=1e-15)
    assert np.allclose(j2, [-0.17210943340705884, 0.0, 1.2207820523028392], atol=1e-15)
    assert np.allclose(j3, [0.15775588001646562, 0.0, 1.0075893577750952], atol=1e-15)
    # the only numeric recorded in the reference: manipulator_3DoF.py:223 is the IK answer for (0.6, 0, 0.1)
    (ex, ez), _, _ = M.arm_fk(np.array([0.42323673, -1.39921683, 1.15256477]))
    assert abs(ex - 0.6) < 1e-8 and abs(ez - 0.1) < 1e-8


def test_compact_fk_equals_reference_expression_order():
    rng = np.random.default_rng(0)
    q = rng.uniform(-4, 4, (2000, 3))
    a, b = M.arm_fk(q), M.arm_fk_compact(q)
    for i in range(3):
        assert np.abs(np.array(a[i]) - np.array(b[i])).max() < 1e-15 * 8


def test_f_kinematics_reference_step():
    x = np.array([1.0, 2.0, 0.3, 0.5, -0.2, 0.1, 0.1, -0.2, 0.3])
    u = np.array([0.7, -0.4, 0.1, 0.2, -0.3])
    dt = 0.1
    exp = [1.05, 1.98, 0.31, 0.5 + dt * (0.7 * np.cos(0.3) + 0.2 * 0.1), -0.2 + dt * (0.7 * np.sin(0.3) + 0.5 * 0.1),
           0.1 - 0.04, 0.11, -0.18, 0.27]
    assert np.allclose(M.f_kinematics(x, u, dt), exp, atol=1e-15)


def test_rows_sympy_consistency():
    """same restated expressions evaluated symbolically (sympy) and numerically agree"""
    import sympy as sp
    xs = sp.symbols("x0:9")
    circ = [(2.5, 1.0, 0.6)]
    planes = [((4.577, 5.0, 1.209), (0.0, 0.0, -1.0)), ((4.577, 5.0, 1.209), (-1.0, 0.0, 0.0))]
    exprs = M.circle_rows(xs, circ, sp) + M.self_collision_rows(xs, sp) + [c for row in M.plane_margins(xs, planes, sp) for c in row]
    f = sp.lambdify(xs, exprs, "numpy")
    rng = np.random.default_rng(1)
    for _ in range(5):
        xv = rng.uniform(-1, 1, 9)
        num = (M.circle_rows(xv, circ) + M.self_collision_rows(xv) + [c for row in M.plane_margins(xv, planes) for c in row])
        assert np.allclose(f(*xv), num, rtol=1e-13, atol=1e-14)


def test_angle_diff():
    assert abs(M.angle_diff(-3.14, 3.14) - (2 * np.pi - 6.28)) < 1e-12
    assert abs(M.angle_diff(0.3, 0.1) - 0.2) < 1e-15
    assert abs(M.angle_diff(3.0, -3.0) - (6.0 - 2 * np.pi)) < 1e-12
