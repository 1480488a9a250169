"""CPU oracle -- robot model functions of the whole-body MPC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
use it, and there only as the checker.

PARITY UNPINNED: the reference delegates all arithmetic to the third-party wheel
``casadi==3.6.4`` (``requirements.txt:9``), which is not installable in this image, and the
reference ships no tests or golden vectors.  This module restates the
reference's *expressions* term by term (citations below, relative to /root/reference) and is
pinned by (i) the symbolic DH derivation ``utils/dh_to_kinematics.py`` (runs here with sympy),
(ii) the known-answer values listed in SURVEY.md section 8(c), see ``tests/test_oracle_model.py``.

Every function takes an optional math namespace ``m`` (``numpy`` by default, ``sympy`` for the
symbolic cross-checks) so the same restated expression is used numerically and symbolically.
Vector arguments are indexable (last axis = component when ``m`` is numpy and the input is a
batch ``[..., 9]``).
"""
import numpy as np

# ---- constants -----------------------------------------------------------------------------
# robot_models/manipulator_3DoF.py:18-22  (classical-DH link lengths of the reduced Panda)
A2, A3, A5, A6, A7 = 0.316, 0.0825, 0.384, 0.088, 0.107
# robot_models/mobile_manipulator.py:14-15
BASELINK2JOINT1_X = -0.007
BASELINK2JOINT1_Z = 0.606 + 0.333
# robot_models/base.py:15
BASE_RADIUS = 0.4
# controllers/mpc_wholebody_qref.py:43-44
ENDPOINT_SELF_COLLISION_RADIUS = 0.05
OBSTACLE_EXPAND_DIST = 0.03

NX, NU = 9, 5


def _c(x, i):
    """component i along the last axis (works for lists, sympy Matrices and numpy batches)."""
    if isinstance(x, np.ndarray):
        return x[..., i]
    return x[i]


# ---- dynamics ------------------------------------------------------------------------------
def f_kinematics(x, u, dt, m=np):
    """One explicit-Euler step of the 9-state model.

    robot_models/mobile_manipulator.py:57-75 -> robot_models/base.py:17-31 (base, unicycle with
    acceleration inputs) + robot_models/manipulator_3DoF.py:189-191 (q += q_dot*dt).
    Returns the 9 components as a list (numpy: stacked on the last axis).
    """
    x0, x1, x2, x3, x4, x5 = (_c(x, i) for i in range(6))
    u0, u1 = _c(u, 0), _c(u, 1)
    nxt = [
        x0 + dt * x3,                                   # base.py:20
        x1 + dt * x4,                                   # base.py:21
        x2 + dt * x5,                                   # base.py:22
        x3 + dt * (u0 * m.cos(x2) - x4 * x5),           # base.py:23
        x4 + dt * (u0 * m.sin(x2) + x3 * x5),           # base.py:24
        x5 + dt * u1,                                   # base.py:25
        _c(x, 6) + _c(u, 2) * dt,                       # manipulator_3DoF.py:190
        _c(x, 7) + _c(u, 3) * dt,
        _c(x, 8) + _c(u, 4) * dt,
    ]
    if m is np:
        return np.stack(np.broadcast_arrays(*nxt), axis=-1)
    return nxt


# ---- forward kinematics --------------------------------------------------------------------
def arm_fk(q, m=np):
    """Planar FK of the 3-DoF arm in the arm frame; returns (endpoint, joint2, joint3), each a
    pair (x, z) -- the y component is identically 0 (manipulator_3DoF.py:33,47,66).

    Expressions copied term by term from the translation columns of the homogeneous transforms
    at robot_models/manipulator_3DoF.py:30,32 (joint 2), :44,:51 (joint 3), :63,:70 (endpoint).
    """
    q1, q2, q3 = _c(q, 0), _c(q, 1), _c(q, 2)
    s1, c1, s2, c2, s3, c3 = m.sin(q1), m.cos(q1), m.sin(q2), m.cos(q2), m.sin(q3), m.cos(q3)
    a2, a3, a5, a6, a7 = A2, A3, A5, A6, A7
    j2x = a2 * s1 + a3 * c1                                                         # :30
    j2z = a2 * c1 - a3 * s1                                                         # :32
    j3x = a2 * s1 - a3 * s1 * s2 - a3 * c1 * c2 + a3 * c1 + a5 * (s1 * c2 - s2 * c1)   # :44
    j3z = a2 * c1 + a3 * s1 * c2 - a3 * s1 - a3 * s2 * c1 + a5 * (s1 * s2 + c1 * c2)   # :51
    ex = (a2 * s1 - a3 * s1 * s2 - a3 * c1 * c2 + a3 * c1 + a5 * (s1 * c2 - s2 * c1)
          - a6 * (-s1 * s2 - c1 * c2) * c3 + a6 * (s1 * c2 - s2 * c1) * s3
          - a7 * ((-s1 * s2 - c1 * c2) * s3 + (s1 * c2 - s2 * c1) * c3))             # :63
    ez = (a2 * c1 + a3 * s1 * c2 - a3 * s1 - a3 * s2 * c1 + a5 * (s1 * s2 + c1 * c2)
          + a6 * (s1 * s2 + c1 * c2) * s3 - a6 * (s1 * c2 - s2 * c1) * c3
          - a7 * ((s1 * s2 + c1 * c2) * c3 + (s1 * c2 - s2 * c1) * s3))              # :70
    return (ex, ez), (j2x, j2z), (j3x, j3z)


def forward_transformation(state, m=np):
    """World-frame endpoint pose (x, y, z, psi), joint-2 and joint-3 positions (x, y, z).

    robot_models/mobile_manipulator.py:17-55 (spelt ``forward_tranformation`` there).
    Returned as three lists of scalars/arrays.
    """
    x, y, psi = _c(state, 0), _c(state, 1), _c(state, 2)
    q = [_c(state, 6), _c(state, 7), _c(state, 8)]
    (ex, ez), (j2x, j2z), (j3x, j3z) = arm_fk(q, m)
    cp, sp = m.cos(psi), m.sin(psi)
    bx, bz = BASELINK2JOINT1_X, BASELINK2JOINT1_Z
    pose_endpoint = [x + (ex + bx) * cp, y + (ex + bx) * sp, 0 + ez + bz, psi]       # :36-41
    pos_joint_2 = [x + (j2x + bx) * cp, y + (j2x + bx) * sp, 0 + j2z + bz]           # :43-47
    pos_joint_3 = [x + (j3x + bx) * cp, y + (j3x + bx) * sp, 0 + j3z + bz]           # :49-53
    return pose_endpoint, pos_joint_2, pos_joint_3


def body_points(state, m=np):
    """The 6 manipulator body points and 4 self-collision check points of one stage.

    controllers/mpc_wholebody_qref.py:213-219.  NB the check points are the *world origin* and
    *half the world position* of joint 2 etc. -- artefacts kept for parity.
    Each point is a list [x, y, z].
    """
    pe, j2, j3 = forward_transformation(state, m)
    e = pe[0:3]
    half = lambda p: [c / 2 for c in p]
    mid = lambda a, b: [(ca + cb) / 2 for ca, cb in zip(a, b)]
    positions = [half(j2), j2, mid(j2, j3), j3, mid(j3, e), e]                      # :216-217
    zero = 0 * e[0]
    checks = [[zero, zero, zero], half(j2), j2, mid(j2, j3)]                         # :219
    return positions, checks


# ---- inequality rows -----------------------------------------------------------------------
def circle_rows(state, circles, m=np):
    """g_i = (r_i + base_radius) - sqrt((x-ox_i)^2 + (y-oy_i)^2)   (<= s_k)

    controllers/mpc_wholebody_qref.py:49-54 (obsAvoid), used at :208-209 and :248-249.
    ``circles`` is a sequence of (ox, oy, r).
    """
    x, y = _c(state, 0), _c(state, 1)
    return [(r + BASE_RADIUS) - m.sqrt((x - ox) ** 2 + (y - oy) ** 2) + 0.0 for (ox, oy, r) in circles]


def self_collision_rows(state, m=np):
    """0.05 - ||check_i - endpoint||_2   (< s_k), i = 0..3.  mpc_wholebody_qref.py:219-222."""
    positions, checks = body_points(state, m)
    e = positions[-1]
    out = []
    for chk in checks:
        d = [a - b for a, b in zip(chk, e)]
        out.append(ENDPOINT_SELF_COLLISION_RADIUS - m.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2))
    return out


def plane_margins(state, planes, m=np):
    """c[i][j] = n_j . ((p_j - 0.03 n_j) - pos_i)  for the 6 body points i and planes j.

    controllers/mpc_wholebody_qref.py:76-80.  ``planes`` is a sequence of (point[3], normal[3]).
    The constraint built from these is  -max_j(...) < s_k  (:82-89); the max and the
    stale-column quirk are assembled in oracle/nlp.py.
    """
    positions, _ = body_points(state, m)
    c = []
    for pos in positions:
        row = []
        for (pt, nrm) in planes:
            pe = [pt[a] - OBSTACLE_EXPAND_DIST * nrm[a] for a in range(3)]          # :78
            row.append(sum(nrm[a] * (pe[a] - pos[a]) for a in range(3)))            # :79-80
        c.append(row)
    return c


# ---- misc helpers of the class -------------------------------------------------------------
def angle_diff(a, b):
    """a-b wrapped to the closest representative.  mpc_wholebody_qref.py:92-117 (numeric form;
    casadi fmod == C fmod, sign of the dividend)."""
    a = np.fmod(a + np.pi, 2 * np.pi) - np.pi
    b = np.fmod(b + np.pi, 2 * np.pi) - np.pi
    d = a - b
    if a * b >= 0:
        return d
    if a > b:
        return d if d <= np.pi else d - 2 * np.pi
    return d if d > -np.pi else d + 2 * np.pi


# ---- compact (theta-chain) FK used by the solvers; verified against arm_fk in tests ---------
def arm_fk_compact(q, m=np):
    """theta1=q1, theta2=q1-q2, theta3=q1-q2-q3; segments v1,v2,v3."""
    t1 = _c(q, 0)
    t2 = t1 - _c(q, 1)
    t3 = t2 - _c(q, 2)
    v1 = (A2 * m.sin(t1) + A3 * m.cos(t1), A2 * m.cos(t1) - A3 * m.sin(t1))
    v2 = (-A3 * m.cos(t2) + A5 * m.sin(t2), A3 * m.sin(t2) + A5 * m.cos(t2))
    v3 = (A6 * m.cos(t3) - A7 * m.sin(t3), -A6 * m.sin(t3) - A7 * m.cos(t3))
    j2 = v1
    j3 = (v1[0] + v2[0], v1[1] + v2[1])
    e = (j3[0] + v3[0], j3[1] + v3[1])
    return e, j2, j3
