"""The oracle against the REFERENCE'S OWN CODE (CPU tier).

tests/golden/ref_*.npz were written by tests/golden/make_ref_rows.py, which runs the unmodified
/root/reference/robot_models/*.py and controllers/mpc_wholebody_qref.py::MPCWholeBody.reset() (:142-285) on numbers
(tests/refshim/casadi.py evaluates the casadi calls eagerly).  Here oracle/model.py and oracle/nlp.py must reproduce
every one of those rows -- dynamics, boxes, circles, self-collision rows with the slack the reference wrote, plane
rows with their stale columns -- and the cost, to 1e-12.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

from oracle import model as M
from oracle import nlp as ONLP

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
from ref_points import random_points, input_checksums  # noqa: E402

T_DYN, T_X0, T_BOXU, T_BOXX, T_BOXDU, T_CIRC, T_SELF, T_PLANE = range(8)
CASES = ["s1", "s2", "s1_N10", "c3", "p1", "p0", "s1_manip"]
RTOL = 1e-12


def close(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    sc = np.maximum(1.0, np.abs(b)) if scale is None else scale
    return np.abs(a - b) <= RTOL * sc


def test_oracle_model_matches_reference_model_code():
    """f_kinematics (mobile_manipulator.py:57-75), forward_tranformation (:17-55), obsAvoid (:49-54), the self-collision
    rows (:219-222) and the plane margins (:76-80), 1,000 random states, values produced by the reference's code."""
    g = np.load(os.path.join(GOLD, "ref_model_values.npz"))
    x, u = g["x"], g["u"]
    assert close(M.f_kinematics(x, u, 0.1), g["f"]).all()
    pe, j2, j3 = M.forward_transformation(x)
    fk = np.stack(np.broadcast_arrays(*(pe + j2 + j3)), axis=-1)
    assert close(fk, g["fk"]).all()
    rows = np.stack(M.circle_rows(x, g["circles"]), axis=-1)
    assert close(rows, g["circle_rows"]).all()
    assert close(np.stack(M.self_collision_rows(x), axis=-1), g["self_rows"]).all()
    for nm in ("s1", "s2"):
        planes = [(p[:3], p[3:]) for p in g["planes_" + nm]]
        c = M.plane_margins(x, planes)
        c = np.stack([np.stack(np.broadcast_arrays(*row), axis=-1) for row in c], axis=-2)
        assert close(c, g["margins_" + nm]).all()


def _nlp(g, pts, m, mode):
    kw = {}
    if "w_Q" in g.files:
        kw = dict(Qd=np.diag(g["w_Q"]), Pd=np.diag(g["w_P"]))
    planes = [p for p in g["planes"]]
    return ONLP.NLP(int(g["N"]), float(g["dt"]), pts["X"][m, 0], pts["X_ref"][m], pts["U_ref"][m], pts["U_last"][m],
                    g["circles"].reshape(-1, 3), planes, mode=mode, **kw)


def oracle_rows_in_reference_order(nlp, X, U, s, tags, vacuous, mode):
    """oracle/nlp.py's eq(), ineq() and box rows laid out in the reference's issue order (NaN where the oracle has no row)."""
    N = nlp.N
    npl = len(nlp.planes)
    w = nlp.pack(X, U, s)
    e = nlp.eq(w).reshape(N, 9)
    gi = nlp.ineq(w, with_boxes=False)
    # index of (type, k, i, j) in ineq()'s order: circles, self-collision, planes (oracle/nlp.py:ineq)
    pos, n = {}, 0
    ncirc = nlp.circles.shape[1]
    for k in range(N + 1):
        for i in range(ncirc):
            pos[(T_CIRC, k, i, -1)] = n; n += 1
    for k in range(N + 1):
        for m in range(4):
            pos[(T_SELF, k, m, -1)] = n; n += 1
    if npl:
        for k in range(N + 1):
            for i in range(6):
                for j in ([npl - 1] if mode == "clean" else range(npl)):
                    if j < npl - 1 and k == 0:
                        continue
                    pos[(T_PLANE, k, i, j)] = n; n += 1
    assert n == gi.size
    out = np.full(len(tags), np.nan)
    for r, (t, k, i, j, ks) in enumerate(tags):
        if t == T_DYN:
            out[r] = -e[k, i]                       # reference: X[k+1] - f ; oracle: f - X[k+1]
        elif t == T_X0:
            out[r] = 0.0                            # X[0] == X_init is substituted in the oracle
        elif t == T_BOXU:
            out[r] = U[k, i]
        elif t == T_BOXX:
            out[r] = X[k, i]
        elif t == T_BOXDU:
            out[r] = U[k, i] - nlp.u_last[k, i]
        elif (t, k, i, j) in pos and not vacuous[r]:
            out[r] = gi[pos[(t, k, i, j)]]
    return out


@pytest.mark.parametrize("case", CASES)
def test_oracle_nlp_rows_match_reference_reset(case):
    g = np.load(os.path.join(GOLD, "ref_rows_%s.npz" % case))
    N, npl = int(g["N"]), g["planes"].shape[0]
    tags, vac, wts = g["tags"], g["vacuous"], g["weights"]
    Mf, Ms = int(g["M_full"]), int(g["M_sum"])
    pts = random_points(np.random.default_rng(int(g["seed"])), Mf + Ms, N, npl)
    assert np.array_equal(input_checksums(pts), g["input_checksums"]), "NumPy's Generator stream differs from the one the goldens were made with"
    # what the reference's row list looks like (now read off the reference's own output)
    term_self = tags[(tags[:, 0] == T_SELF) & (tags[:, 1] == N)]
    assert (term_self[:, 4] == N - 1).all()                                   # quirk 3: the leaked loop variable
    other = tags[(tags[:, 0] >= T_CIRC) & ~((tags[:, 0] == T_SELF) & (tags[:, 1] == N))]
    assert (other[:, 4] == other[:, 1]).all()                                 # every other row is bounded by its own stage's slack
    vt = tags[vac]
    assert len(vt) == (6 * (npl - 1) if npl > 1 else 0)
    assert ((vt[:, 0] == T_PLANE) & (vt[:, 1] == 0) & (vt[:, 3] < npl - 1)).all()   # quirk 2: exactly the k = 0, j < n_pl - 1 rows
    # the box bounds the reference wrote are the oracle's
    n0 = _nlp(g, pts, 0, "reference")
    for t, lim in ((T_BOXU, n0.ulim), (T_BOXX, n0.xlim), (T_BOXDU, n0.dulim)):
        sel = tags[:, 0] == t
        assert np.array_equal(g["box_lo"][sel], lim[0][tags[sel, 2]]) and np.array_equal(g["box_hi"][sel], lim[1][tags[sel, 2]])
    sums = np.zeros((Ms, 8))
    scale = np.zeros((Ms, 8))
    for m in range(Mf + Ms):
        nlp = _nlp(g, pts, m, "reference")
        assert np.array_equal(nlp.x_init, pts["X"][m, 0])
        X, U, s = pts["X"][m], pts["U"][m], pts["s"][m, :, 0]
        o = oracle_rows_in_reference_order(nlp, X, U, s, tags, vac, "reference")
        assert np.isfinite(o[~vac]).all()
        assert close(nlp.cost(nlp.pack(X, U, s)), g["cost"][m], scale=abs(g["cost"][m])), (case, m)
        if m < Mf:
            bad = ~close(o[~vac], g["rows_full"][m][~vac])
            assert not bad.any(), (case, m, tags[~vac][bad][:5])
        else:
            for t in range(8):
                sel = (tags[:, 0] == t) & ~vac
                sums[m - Mf, t] = wts[sel] @ o[sel]
                scale[m - Mf, t] = max(1.0, wts[sel] @ np.abs(o[sel]))
    assert close(sums, g["row_sums"], scale=scale).all()
    # the stage-separable variant keeps the proper rows (j = n_pl - 1) and moves the terminal self-collision rows to s[N]
    for m in range(min(Mf, 4)):
        nlp = _nlp(g, pts, m, "clean")
        X, U, s = pts["X"][m], pts["U"][m], pts["s"][m, :, 0]
        o = oracle_rows_in_reference_order(nlp, X, U, s, tags, vac, "clean")
        ref = g["rows_full"][m].copy()
        ts = (tags[:, 0] == T_SELF) & (tags[:, 1] == N)
        ref[ts] += s[N - 1] - s[N]
        keep = np.isfinite(o) & ~vac
        assert keep.sum() == (~vac).sum() - 6 * max(npl - 1, 0) * N
        assert close(o[keep], ref[keep]).all()


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference checkout only exists in the authoring container")
def test_goldens_are_what_the_reference_code_produces_now():
    """Re-runs the reference's reset() under the stand-in and compares with the committed file (scenario 2)."""
    spec = importlib.util.spec_from_file_location("make_ref_rows", os.path.join(GOLD, "make_ref_rows.py"))
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    try:
        mk = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mk)
        g = np.load(os.path.join(GOLD, "ref_rows_s2.npz"))
        pts = random_points(np.random.default_rng(int(g["seed"])), int(g["M_full"]) + int(g["M_sum"]), 20, 2)
        ctrl = mk.run_reset(20, mk.DEMO_CIRCLES, mk.SCEN2, pts)
        vals, tags, _, _ = mk.flatten(ctrl, pts, 3, 2)
        assert np.array_equal(tags, g["tags"])
        assert np.array_equal(vals[:int(g["M_full"])][:, ~g["vacuous"]], g["rows_full"][:, ~g["vacuous"]])
        assert np.array_equal(np.broadcast_to(ctrl.cost.v, (vals.shape[0], 1, 1)).reshape(-1), g["cost"])
    finally:
        sys.path[:] = saved_path
        for k in set(sys.modules) - saved_mods:   # the stand-in `casadi` and the reference packages must not leak into other tests
            del sys.modules[k]


def test_oracle_nlp_base_rows_match_reference_mpcbase_reset():
    """oracle/nlp_base.py against the rows and the cost that the reference's own
    controllers/mpc_base.py::MPCBase.reset() (:114-189) produced (tests/golden/ref_rows_base.npz): dynamics, boxes, circle rows
    with their slack, and the cost with its angleDiff yaw error (a yaw weight is set so that it shows)."""
    from ref_points import random_points_base
    from oracle.nlp_base import NLPBase
    g = np.load(os.path.join(GOLD, "ref_rows_base.npz"))
    N, Mf, Ms = int(g["N"]), int(g["M_full"]), int(g["M_sum"])
    tags, wts = g["tags"], g["weights"]
    pts = random_points_base(np.random.default_rng(int(g["seed"])), Mf + Ms, N)
    assert np.array_equal(input_checksums(pts), g["input_checksums"])
    sums, scale = np.zeros((Ms, 8)), np.zeros((Ms, 8))
    for m in range(Mf + Ms):
        P = NLPBase(N, float(g["dt"]), pts["X"][m, 0], pts["X_ref"][m], pts["U_ref"][m], g["circles"], Qd=g["Qd"], Pd=g["Pd"])
        X, U, s = pts["X"][m], pts["U"][m], pts["s"][m, :, 0]
        w = P.pack(X, U, s)
        e = P.eq(w).reshape(N, 6)
        gi = P.ineq(w, with_boxes=False).reshape(N + 1, -1)
        o = np.empty(len(tags))
        for r, (t, k, i, j, ks) in enumerate(tags):
            o[r] = {T_DYN: lambda: -e[k, i], T_X0: lambda: 0.0, T_BOXU: lambda: U[k, i], T_BOXX: lambda: X[k, i],
                    T_BOXDU: lambda: X[k, 3 + i], T_CIRC: lambda: gi[k, i]}[t]()
            assert t != T_CIRC or ks == k
        assert close(P.cost(w), g["cost"][m], scale=abs(g["cost"][m])), m
        if m < Mf:
            assert close(o, g["rows_full"][m]).all(), m
        else:
            for t in range(8):
                sel = tags[:, 0] == t
                sums[m - Mf, t] = wts[sel] @ o[sel]; scale[m - Mf, t] = max(1.0, wts[sel] @ np.abs(o[sel]))
    assert close(sums, g["row_sums"], scale=scale).all()
