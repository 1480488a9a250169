#!/usr/bin/env python
"""Golden vectors produced by the UNMODIFIED reference code, executed in this container.

    python tests/golden/make_ref_rows.py            (needs /root/reference; writes tests/golden/ref_*.npz)

``tests/refshim/casadi.py`` is a numeric stand-in for the casadi calls the reference makes while it BUILDS
its NLP.  With it first on sys.path this script imports /root/reference/robot_models/*.py and
/root/reference/controllers/mpc_wholebody_qref.py as they are and lets ``MPCWholeBody.reset()`` (:142-285)
run on numeric (X, U, s, parameters): what comes out is the reference's own row list -- every
``opti.subject_to`` in issue order with the stale ``self.constr`` columns (:76-89), the free ``constr``
variables (:156) and the leaked loop variable of the terminal self-collision rows (:263-265) -- and its cost
(:192-201, 227, 240-242, 270).  The files written here pin oracle/model.py, oracle/nlp.py and the device
model (mmpc_eval_model) to reference code instead of to a restatement.

Files:
  ref_model_values.npz   1,000 random (x, u): f_kinematics, forward_tranformation, obsAvoid rows, the four
                         self-collision rows and the plane margins c[i][j] of demo scenarios 1 and 2
  ref_rows_<case>.npz    per case: inputs of 16 + 1,024 random points, the tag table of the reference's rows
                         (type, stage, i, j, slack index -- the slack index is READ OFF the right-hand side the
                         reference wrote, not assumed), all row values at the first 16 points, and for the
                         other 1,024 points the cost and one weighted checksum of all rows per row type.  The random inputs are
                         regenerated from the stored seed (tests/golden/ref_points.py; checksums stored)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MMPC_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "tests", "refshim"), REF]

sys.path.insert(0, HERE)
from ref_points import random_points, random_points_base, input_checksums  # noqa: E402
import casadi as ca  # noqa: E402  (the stand-in)
from controllers.mpc_wholebody_qref import MPCWholeBody  # noqa: E402  (reference, unmodified)
from robot_models.mobile_manipulator import MobileManipulator  # noqa: E402
from robot_models.obstacles import Obstacles  # noqa: E402

assert ca.__file__.startswith(os.path.join(ROOT, "tests", "refshim"))
assert MPCWholeBody.__module__ == "controllers.mpc_wholebody_qref" and REF in sys.modules[MPCWholeBody.__module__].__file__

PI = np.pi
# row types of the tag table
T_DYN, T_X0, T_BOXU, T_BOXX, T_BOXDU, T_CIRC, T_SELF, T_PLANE = range(8)

# demo_wholebody_qref.py:18-44, typed as the demo types them
P1 = np.array([5.007 - 0.43, 5, 0.27 + 0.606 + 0.333])
SCEN1 = [(P1, np.array([[0, 0, -1]])), (P1, np.array([[-1, 0, 0]])), (P1, np.array([[0, 1, 0]]))]
P2 = np.array([2.5, 2, 0.35 + 0.606 + 0.333])
SCEN2 = [(P2, np.array([[1 / np.sqrt(2), 0, 1 / np.sqrt(2)]])), (P2, np.array([[-1 / np.sqrt(2), 0, 1 / np.sqrt(2)]]))]
DEMO_CIRCLES = [(2.5, 3.0, 0.6), (2.5, 1.0, 0.6), (5 - 0.6, 5, 0.1)]


def run_reset(N, circles, planes, pts, dt=0.1, weights=None):
    """The reference's reset() on the fed numbers; returns the controller."""
    ca.Opti.FEED = dict(variable=[pts["X"], pts["U"], pts["s"], pts["free"]],
                        parameter=[pts["U_last"], pts["X_init"], pts["X_ref"], pts["U_ref"]])
    robot = MobileManipulator(dt)
    obs = [Obstacles(*c) for c in circles]
    ctrl = MPCWholeBody(robot, obs, planes, N=N, **(weights or {}))
    ca.Opti.FEED = None
    return ctrl


def flatten(ctrl, pts, n_obs, npl):
    """Reference rows in issue order -> values [M, R], tags [R, 5] = (type, stage, i, j, slack index), box bounds."""
    N = ctrl.N
    cons = ctrl.opti.constraints
    M = pts["X"].shape[0]
    s = pts["s"][:, :, 0]
    vals, tags, blo, bhi = [], [], [], []
    it = iter(cons)

    def bc(a):
        a = np.asarray(a)
        return np.broadcast_to(a, (M,) + a.shape[-2:]).reshape(M, -1)

    def take_eq(typ, k):
        c = next(it)
        assert c.op == "==" and c.lhs.shape == (1, 9), (c.op, c.lhs.shape)
        v = bc(c.lhs.v - c.rhs.v)
        for i in range(9):
            vals.append(v[:, i]); tags.append((typ, k, i, -1, -1)); blo.append(np.nan); bhi.append(np.nan)

    def take_box(typ, k, n):
        c = next(it)
        assert c.op == "bounded" and c.mid.shape == (1, n), (c.op, c.mid.shape)
        v, lo, hi = bc(c.mid.v), bc(c.lhs.v), bc(c.rhs.v)
        for i in range(n):
            vals.append(v[:, i]); tags.append((typ, k, i, -1, -1)); blo.append(lo[0, i]); bhi.append(hi[0, i])

    def take_row(typ, k, i, j, op):
        c = next(it)
        assert c.op == op and c.lhs.shape == (1, 1), (c.op, op, c.lhs.shape)
        rhs = bc(c.rhs.v)[:, 0]
        ks = [q for q in range(N + 1) if np.array_equal(rhs, s[:, q])]   # which slack did the reference write?
        assert len(ks) == 1, ks
        vals.append(bc(c.lhs.v)[:, 0] - rhs); tags.append((typ, k, i, j, ks[0])); blo.append(np.nan); bhi.append(np.nan)

    def take_stage_rows(k):
        for i in range(n_obs):
            take_row(T_CIRC, k, i, -1, "<=")          # :208-209 / :248-249
        for m in range(4):
            take_row(T_SELF, k, m, -1, "<")           # :220-222 / :263-265
        for i in range(6):
            for j in range(npl):
                take_row(T_PLANE, k, i, j, "<")       # :76-89

    for k in range(N):
        take_eq(T_DYN, k)                              # :180
        take_box(T_BOXU, k, 5)                         # :203
        take_box(T_BOXX, k, 9)                         # :204
        take_box(T_BOXDU, k, 5)                        # :205
        take_stage_rows(k)
    take_eq(T_X0, 0)                                   # :244
    take_box(T_BOXX, N, 9)                             # :245
    take_stage_rows(N)
    assert next(it, None) is None, "unconsumed reference constraints"
    return np.stack(vals, axis=1), np.array(tags, dtype=np.int32), np.array(blo), np.array(bhi)


def make_case(name, N, circles, planes, seed, M_full=16, M_sum=1024, weights=None):
    rng = np.random.default_rng(seed)
    npl = len(planes)
    pts = random_points(rng, M_full + M_sum, N, npl)
    ctrl = run_reset(N, circles, planes, pts, weights=weights)
    vals, tags, blo, bhi = flatten(ctrl, pts, len(circles), npl)
    cost = np.broadcast_to(ctrl.cost.v, (M_full + M_sum, 1, 1)).reshape(-1)
    # a second pass with other values of the free `constr` variables: the rows that move are the vacuous ones (quirk 2)
    pts2 = dict(pts); pts2["free"] = rng.normal(size=pts["free"].shape)
    vals2, _, _, _ = flatten(run_reset(N, circles, planes, pts2, weights=weights), pts2, len(circles), npl)
    vacuous = np.any(vals != vals2, axis=0)
    wts = rng.uniform(0.5, 1.5, size=vals.shape[1])
    sums = np.stack([np.where((tags[:, 0] == t) & ~vacuous, wts, 0.0) @ vals[M_full:].T for t in range(8)], axis=1)
    pl = np.array([np.hstack([np.asarray(p, float).reshape(3), np.asarray(n, float).reshape(3)]) for p, n in planes]).reshape(npl, 6)
    out = dict(N=N, dt=0.1, circles=np.array(circles, float).reshape(-1, 3), planes=pl, tags=tags, box_lo=blo, box_hi=bhi,
               vacuous=vacuous, weights=wts, M_full=M_full, M_sum=M_sum, seed=seed, rows_full=vals[:M_full], cost=cost, row_sums=sums,
               # the inputs are regenerated from the seed by tests/golden/ref_points.py; their checksums guard the RNG stream
               input_checksums=np.array([float(np.sum(pts[k] * np.cos(np.arange(pts[k].size).reshape(pts[k].shape)))) for k in sorted(pts)]))
    if weights:
        out.update({"w_" + k: np.asarray(v, float) for k, v in weights.items()})
    np.savez_compressed(os.path.join(HERE, "ref_rows_%s.npz" % name), **out)
    print("%-8s N=%d n_obs=%d n_pl=%d rows=%d vacuous=%d slack-of-terminal-self-rows=%s" % (
        name, N, len(circles), npl, vals.shape[1], int(vacuous.sum()),
        sorted(set(tags[(tags[:, 0] == T_SELF) & (tags[:, 1] == N), 4].tolist()))))


def make_model_values(M=1000, seed=11):
    rng = np.random.default_rng(seed)
    pts = random_points(rng, M, 1, 3)
    out = dict(x=pts["X"][:, 1, :], u=pts["U"][:, 0, :])
    robot = MobileManipulator(0.1)
    x, u = ca.NM(pts["X"][:, 1:2, :].copy()), ca.NM(pts["U"][:, 0:1, :].copy())
    out["f"] = np.asarray(robot.f_kinematics(x, u)).reshape(M, 9)                      # mobile_manipulator.py:57-75
    pe, j2, j3 = robot.forward_tranformation(x)                                          # :17-55
    out["fk"] = np.concatenate([np.asarray(a).reshape(M, -1) for a in (pe, j2, j3)], axis=1)
    rc = rng.uniform([0.5, 0.5, 0.1], [5.5, 5.5, 0.6], size=(16, 3))
    obs = [Obstacles(*c) for c in list(DEMO_CIRCLES) + [tuple(r) for r in rc]]
    holder = types.SimpleNamespace(base_radius=robot.base.base_radius())
    out["circles"] = np.array([[o.x, o.y, o.radius] for o in obs])
    out["circle_rows"] = np.concatenate([np.asarray(g).reshape(M, 1) for g in MPCWholeBody.obsAvoid(holder, obs, x)], axis=1)  # :49-54
    # self-collision rows (:219-222) and plane margins c[i][j] (:76-80) are inline in reset(): a horizon-1 reset whose
    # terminal stage sits at x leaves them in the terminal rows (value + slack) and in ctrl.constr
    for nm, planes in (("s1", SCEN1), ("s2", SCEN2)):
        pts_n = dict(pts); pts_n["free"] = pts["free"][:, :, :len(planes)]
        ctrl = run_reset(1, [], planes, pts_n)
        vals, tags, _, _ = flatten(ctrl, pts_n, 0, len(planes))
        sel = (tags[:, 0] == T_SELF) & (tags[:, 1] == 1)
        sl = pts["s"][:, tags[sel, 4], 0]
        out["self_rows"] = vals[:, sel] + sl
        out["margins_" + nm] = np.broadcast_to(ctrl.constr.v, (M, 6, len(planes))).copy()
        out["planes_" + nm] = np.array([np.hstack([np.asarray(p, float).reshape(3), np.asarray(n, float).reshape(3)]) for p, n in planes])
    np.savez_compressed(os.path.join(HERE, "ref_model_values.npz"), **out)
    print("model values: %d points" % M)


def make_base_case(M_full=16, M_sum=512, seed=201, N=10):
    """controllers/mpc_base.py::MPCBase.reset() (:114-189) on numbers: SURVEY.md 8(f) row 4."""
    from controllers.mpc_base import MPCBase          # reference, unmodified
    from robot_models.base import Base                # reference, unmodified
    assert REF in sys.modules[MPCBase.__module__].__file__
    rng = np.random.default_rng(seed)
    M = M_full + M_sum
    pts = random_points_base(rng, M, N)
    Q = np.diag([5., 5., 3.0, 0.5, 0.25, 1.])           # a yaw weight, so that angleDiff (:129-133) shows in the cost
    ca.Opti.FEED = dict(variable=[pts["X"], pts["U"], pts["s"]], parameter=[pts["X_init"], pts["X_ref"], pts["U_ref"]])
    ctrl = MPCBase(Base(0.1), [Obstacles(*c) for c in DEMO_CIRCLES], N=N, Q=Q, P=2 * Q)
    ca.Opti.FEED = None
    it = iter(ctrl.opti.constraints)
    s = pts["s"][:, :, 0]
    vals, tags = [], []
    bc = lambda a: np.broadcast_to(np.asarray(a), (M,) + np.asarray(a).shape[-2:]).reshape(M, -1)

    def take(kind, k, n, op, i0=0):
        c = next(it)
        if op == "bounded":
            assert c.op == "bounded" and c.mid.shape == (1, n), (c.op, c.mid.shape)
            v = bc(c.mid.v)
        else:
            assert c.op == op and c.lhs.shape == (1, n), (c.op, c.lhs.shape)
            v = bc(c.lhs.v - c.rhs.v)
        ks = -1
        if kind == T_CIRC:
            rhs = bc(c.rhs.v)[:, 0]
            (ks,) = [q for q in range(N + 1) if np.array_equal(rhs, s[:, q])]
        for i in range(n):
            vals.append(v[:, i]); tags.append((kind, k, i0 + i, -1, ks))

    for k in range(N):
        take(T_DYN, k, 6, "==")            # :128
        take(T_BOXU, k, 2, "bounded")      # :139
        take(T_BOXX, k, 2, "bounded")      # :140  x, y
        take(T_BOXDU, k, 3, "bounded")     # :141  dx, dy, dpsi  (tag BOXDU reused for the velocity box)
        for i in range(3):
            take(T_CIRC, k, 1, "<=", i)    # :142-143
    take(T_X0, 0, 6, "==")                 # :153
    take(T_BOXX, N, 2, "bounded"); take(T_BOXDU, N, 3, "bounded")   # :154-155
    for i in range(3):
        take(T_CIRC, N, 1, "<=", i)        # :156-157
    assert next(it, None) is None
    vals = np.stack(vals, axis=1); tags = np.array(tags, np.int32)
    cost = np.broadcast_to(ctrl.opti.objective.v, (M, 1, 1)).reshape(-1)
    wts = rng.uniform(0.5, 1.5, size=vals.shape[1])
    sums = np.stack([np.where(tags[:, 0] == t, wts, 0.0) @ vals[M_full:].T for t in range(8)], axis=1)
    np.savez_compressed(os.path.join(HERE, "ref_rows_base.npz"), N=N, dt=0.1, circles=np.array(DEMO_CIRCLES, float), Qd=np.diag(Q), Pd=2 * np.diag(Q),
                        tags=tags, weights=wts, M_full=M_full, M_sum=M_sum, seed=seed, rows_full=vals[:M_full], cost=cost, row_sums=sums,
                        input_checksums=input_checksums(pts))
    print("base     N=%d rows=%d (MPCBase, controllers/mpc_base.py)" % (N, vals.shape[1]))
    return pts


def make_pose_case(M_full=16, M_sum=512, seed=301, N=10):
    """controllers/mpc_wholebody.py::MPCWholeBody.reset() (:49-128), the pose-reference sibling (end-point pose in the cost), on
    numbers: SURVEY.md 8(f) row 4."""
    import importlib
    mod = importlib.import_module("controllers.mpc_wholebody")     # reference, unmodified
    assert REF in mod.__file__
    rng = np.random.default_rng(seed)
    M = M_full + M_sum
    pts = random_points(rng, M, N, 0)
    pts = {k: v for k, v in pts.items() if k != "free"}
    pts["X_ref"] = np.ascontiguousarray(pts["X_ref"][:, :, :4])                     # x y z psi of the end point (:66)
    pts["X_ref"][:, :, 2] = rng.uniform(0.6, 1.8, size=(M, N + 1))                   # a reachable height
    Q = np.diag([5., 4., 3., 2.]); P = np.diag([50., 40., 30., 20.])                # distinct weights: every component shows in the cost
    ca.Opti.FEED = dict(variable=[pts["X"], pts["U"], pts["s"]], parameter=[pts["U_last"], pts["X_init"], pts["X_ref"], pts["U_ref"]])
    ctrl = mod.MPCWholeBody(MobileManipulator(0.1), [Obstacles(*c) for c in DEMO_CIRCLES], N=N, Q=Q, P=P)
    ca.Opti.FEED = None
    it = iter(ctrl.opti.constraints)
    s = pts["s"][:, :, 0]
    vals, tags, blo, bhi = [], [], [], []
    bc = lambda a: np.broadcast_to(np.asarray(a), (M,) + np.asarray(a).shape[-2:]).reshape(M, -1)

    def take(kind, k, n, op, i0=0):
        c = next(it)
        lo = hi = None
        if op == "bounded":
            assert c.op == "bounded" and c.mid.shape == (1, n), (c.op, c.mid.shape)
            v, lo, hi = bc(c.mid.v), bc(c.lhs.v), bc(c.rhs.v)
        else:
            assert c.op == op and c.lhs.shape == (1, n), (c.op, c.lhs.shape)
            v = bc(c.lhs.v - c.rhs.v)
        ks = -1
        if kind == T_CIRC:
            rhs = bc(c.rhs.v)[:, 0]
            (ks,) = [q for q in range(N + 1) if np.array_equal(rhs, s[:, q])]
        for i in range(n):
            vals.append(v[:, i]); tags.append((kind, k, i0 + i, -1, ks))
            blo.append(np.nan if lo is None else lo[0, i]); bhi.append(np.nan if hi is None else hi[0, i])

    for k in range(N):
        take(T_DYN, k, 9, "==")            # :76
        take(T_BOXU, k, 5, "bounded")      # :91
        take(T_BOXX, k, 9, "bounded")      # :92
        take(T_BOXDU, k, 5, "bounded")     # :93
        for i in range(3):
            take(T_CIRC, k, 1, "<=", i)    # :96-97
    take(T_X0, 0, 9, "==")                 # :108
    take(T_BOXX, N, 9, "bounded")          # :109
    for i in range(3):
        take(T_CIRC, N, 1, "<=", i)        # :112-113
    assert next(it, None) is None
    vals = np.stack(vals, axis=1); tags = np.array(tags, np.int32)
    cost = np.broadcast_to(ctrl.opti.objective.v, (M, 1, 1)).reshape(-1)
    wts = rng.uniform(0.5, 1.5, size=vals.shape[1])
    sums = np.stack([np.where(tags[:, 0] == t, wts, 0.0) @ vals[M_full:].T for t in range(8)], axis=1)
    np.savez_compressed(os.path.join(HERE, "ref_rows_pose.npz"), N=N, dt=0.1, circles=np.array(DEMO_CIRCLES, float), Qd=np.diag(Q), Pd=np.diag(P),
                        tags=tags, box_lo=np.array(blo), box_hi=np.array(bhi), weights=wts, M_full=M_full, M_sum=M_sum, seed=seed,
                        rows_full=vals[:M_full], cost=cost, row_sums=sums, x_ref_z=pts["X_ref"][:, :, 2], input_checksums=input_checksums(pts))
    print("pose     N=%d rows=%d (pose-reference MPCWholeBody, controllers/mpc_wholebody.py)" % (N, vals.shape[1]))
    return pts


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pose":
        make_pose_case(); sys.exit(0)
    make_base_case()
    make_pose_case()
    make_model_values()
    rng = np.random.default_rng(3)
    c16 = [tuple(r) for r in rng.uniform([0.5, 0.5, 0.1], [5.5, 5.5, 0.6], size=(16, 3))]
    make_case("s1", 20, DEMO_CIRCLES, SCEN1, seed=101)                       # demo scenario 1 (BASELINE config 1)
    make_case("s2", 20, DEMO_CIRCLES, SCEN2, seed=102)                       # demo scenario 2 (config 2): if_else branch :85
    make_case("s1_N10", 10, DEMO_CIRCLES, SCEN1, seed=103, M_sum=256)        # the class's default horizon :11
    make_case("c3", 20, c16, SCEN1, seed=104, M_sum=256)                     # config 3 shape: 16 circles
    make_case("p1", 5, DEMO_CIRCLES[:1], SCEN1[:1], seed=105, M_sum=256)     # one plane: the len == 1 branch :82-83
    make_case("p0", 5, DEMO_CIRCLES, [], seed=106, M_sum=256)                # no planes: obsAvoidConvex never runs :224
    make_case("s1_manip", 20, DEMO_CIRCLES, SCEN1, seed=107, M_sum=256,      # the Interface's 'manipulate' weights :212-215
              weights=dict(Q=np.diag([500, 500, 500, 0, 0, 1, 1, 1, 1.0]), P=np.diag([500, 500, 500, 0, 0, 1, 1, 1, 1.0])))
