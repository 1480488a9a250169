"""SURVEY.md 8(f) row 4, second sibling: the reference's pose-reference whole-body controller (controllers/mpc_wholebody.py: the
tracking cost is on the end-point pose through forward_tranformation).

CPU tier: oracle/nlp_pose.py against the rows and the cost the reference's own reset() (:49-128) produced under the numeric
casadi stand-in (tests/golden/ref_rows_pose.npz, written by tests/golden/make_ref_rows.py::make_pose_case); the oracle's
interior-point solver (MMPC_MODEL_POSEREF) against the restated NLP's KKT conditions.  GPU tier: the CUDA path against the oracle."""
import os
import sys

import numpy as np
import pytest

from mobile_manipulator_mpc_b200 import _abi, scenarios
from oracle.nlp_pose import NLPPose, from_batch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
from ref_points import random_points, input_checksums  # noqa: E402

T_DYN, T_X0, T_BOXU, T_BOXX, T_BOXDU, T_CIRC = range(6)
RTOL = 1e-12


def close(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    sc = np.maximum(1.0, np.abs(b)) if scale is None else scale
    return np.abs(a - b) <= RTOL * sc


def test_oracle_nlp_pose_rows_match_reference_reset():
    g = np.load(os.path.join(GOLD, "ref_rows_pose.npz"))
    N, Mf, Ms = int(g["N"]), int(g["M_full"]), int(g["M_sum"])
    tags, wts = g["tags"], g["weights"]
    rng = np.random.default_rng(int(g["seed"]))
    pts = {k: v for k, v in random_points(rng, Mf + Ms, N, 0).items() if k != "free"}
    pts["X_ref"] = np.ascontiguousarray(pts["X_ref"][:, :, :4])
    pts["X_ref"][:, :, 2] = rng.uniform(0.6, 1.8, size=(Mf + Ms, N + 1))
    assert np.array_equal(input_checksums(pts), g["input_checksums"])
    # the boxes the reference wrote (:16-21)
    P0 = NLPPose(N, 0.1, pts["X"][0, 0], pts["X_ref"][0], pts["U_ref"][0], pts["U_last"][0], g["circles"])
    for t, lim in ((T_BOXU, P0.ulim), (T_BOXX, P0.xlim), (T_BOXDU, P0.dulim)):
        sel = tags[:, 0] == t
        assert np.array_equal(g["box_lo"][sel], lim[0][tags[sel, 2]]) and np.array_equal(g["box_hi"][sel], lim[1][tags[sel, 2]])
    sums, scale = np.zeros((Ms, 8)), np.zeros((Ms, 8))
    for m in range(Mf + Ms):
        P = NLPPose(N, float(g["dt"]), pts["X"][m, 0], pts["X_ref"][m], pts["U_ref"][m], pts["U_last"][m], g["circles"], Qd=g["Qd"], Pd=g["Pd"])
        P.x_init = pts["X"][m, 0]          # the rows are evaluated at the fed point, not at the clipped initial state
        X, U, s = pts["X"][m], pts["U"][m], pts["s"][m, :, 0]
        w = P.pack(X, U, s)
        e = P.eq(w).reshape(N, 9)
        gi = P.ineq(w, with_boxes=False).reshape(N + 1, -1)
        o = np.empty(len(tags))
        for r, (t, k, i, j, ks) in enumerate(tags):
            o[r] = {T_DYN: lambda: -e[k, i], T_X0: lambda: 0.0, T_BOXU: lambda: U[k, i], T_BOXX: lambda: X[k, i],
                    T_BOXDU: lambda: U[k, i] - pts["U_last"][m, k, i], T_CIRC: lambda: gi[k, i]}[t]()
            assert t != T_CIRC or ks == k
        assert close(P.cost(w), g["cost"][m], scale=abs(g["cost"][m])), m
        if m < Mf:
            assert close(o, g["rows_full"][m]).all(), m
        else:
            for t in range(8):
                sel = tags[:, 0] == t
                sums[m - Mf, t] = wts[sel] @ o[sel]; scale[m - Mf, t] = max(1.0, wts[sel] @ np.abs(o[sel]))
    assert close(sums, g["row_sums"], scale=scale).all()


def _oracle_solve(b):
    from oracle import solver
    cfg = solver.config_from_batch(b, _abi.MODE_CLEAN, model=_abi.MODEL_POSEREF)
    for i in range(9):
        cfg.xlim[0][i] = scenarios.POSE_XLIM[0][i]; cfg.xlim[1][i] = scenarios.POSE_XLIM[1][i]
    return solver.solve(b, cfg=cfg, mode=_abi.MODE_CLEAN, threads=os.cpu_count() or 4)


def test_oracle_solver_reaches_kkt_points_of_the_restated_pose_nlp():
    """oracle/mmpc_oracle.c with MMPC_MODEL_POSEREF (exact Hessian of the pose cost) against the dense restatement that is pinned
    to the reference's reset(): feasible, same cost, stationary (non-negative least squares over the active rows' multipliers)."""
    b = scenarios.make_pose_batch(6)
    o = _oracle_solve(b)
    assert (o["status"] == 0).all(), o["status"]
    for i in range(6):
        P = from_batch(b, i)
        w = P.pack(o["X"][i], o["U"][i], o["s"][i])
        assert abs(P.cost(w) - o["cost"][i]) <= 1e-9 * max(1.0, abs(o["cost"][i]))
        assert P.violation(w) <= 1e-8
        assert P.kkt_residual(w) <= 1e-4, (i, P.kkt_residual(w))     # relative to the largest gradient entry (slack cost 2 S s ~ 1e5 when a start lies inside a circle)


def test_dropin_pose_class_surface():
    """controllers/mpc_wholebody.py:6-47 -- constructor, attributes and reset() without a device (the handle is created by the
    first solve)."""
    from mobile_manipulator_mpc_b200.controllers.mpc_wholebody import MPCWholeBody
    from mobile_manipulator_mpc_b200.robot_models import MobileManipulator, Obstacles
    c = MPCWholeBody(MobileManipulator(0.1), [Obstacles(*o) for o in scenarios.DEMO_CIRCLES], N=10)
    assert c.N == 10 and c.base_radius == 0.4 and c._cfg.model == _abi.MODEL_POSEREF and c._cfg.n_pl == 0
    assert list(c.weights["Qd"]) == [5, 5, 5, 5, 0, 0, 0, 0, 0] and list(c.weights["Pd"][:4]) == [50] * 4
    assert np.array_equal(c.xlim, scenarios.POSE_XLIM)
    assert len(c.obsAvoid(c.obstacle_list, np.zeros(9))) == 3


@pytest.mark.gpu
@pytest.mark.parametrize("n_obs,B", [(3, 192), (16, 96)])
def test_gpu_pose_controller_matches_oracle(n_obs, B):
    """The CUDA path (resident kernel, pose build) through the drop-in class against the oracle: north-star tolerances on every
    instance both converge on; then the single-instance call with its warm start."""
    from mobile_manipulator_mpc_b200.controllers.mpc_wholebody import MPCWholeBody
    from mobile_manipulator_mpc_b200.robot_models import MobileManipulator, Obstacles
    b = scenarios.make_pose_batch(B, n_obs=n_obs, seed=12)
    ref = _oracle_solve(b)
    c = MPCWholeBody(MobileManipulator(b["dt"]), [Obstacles(*o) for o in b["circles"][0]], N=b["N"], batch=B)
    out = c.solve_batch(b["x_init"], b["x_ref"][:, :, :4], b["u_ref"], circles=b["circles"])
    assert c._solver.last_solver() == "resident"
    both = (out["status"] == 0) & (ref["status"] == 0)
    assert both.sum() >= B - 2, (np.bincount(out["status"], minlength=6), np.bincount(ref["status"], minlength=6))
    rel = np.abs(out["cost"] - ref["cost"])[both] / np.maximum(1.0, np.abs(ref["cost"][both]))
    du0 = np.abs(out["U"][:, 0] - ref["U"][:, 0]).max(axis=1)[both]
    bad = (rel > 1e-5) | (du0 > 1e-4)
    assert bad.sum() <= 1, (int(bad.sum()), rel.max(), du0.max())      # (a second local optimum of the non-convex pose cost)
    assert (np.abs(out["iters"] - ref["iters"])[both] <= 2).mean() >= 0.9
    for i in np.nonzero(both)[0][:4]:       # the CUDA result is a feasible KKT point of the reference-pinned restatement
        P = from_batch(b, int(i))
        w = P.pack(out["X"][i], out["U"][i], out["s"][i])
        assert P.violation(w) <= 1e-6 and P.kkt_residual(w) <= 1e-4
    # single instance: solve() twice (the second call warm-starts X and U from the first, :138-146, :173-174)
    c1 = MPCWholeBody(MobileManipulator(b["dt"]), [Obstacles(*o) for o in b["circles"][0]], N=b["N"])
    i = int(np.nonzero(both)[0][0])
    u = c1.solve(b["x_init"][i].copy(), b["x_ref"][i, :, :4], b["u_ref"][i])
    assert np.abs(u - ref["U"][i, 0]).max() < 1e-4 and c1.x_guess.shape == (b["N"] + 1, 9)
    xg, ul = c1.x_guess.copy(), c1.u_latest.copy()
    u2 = c1.solve(b["x_init"][i].copy(), b["x_ref"][i, :, :4], b["u_ref"][i])
    assert c1.last_info["status"][0] == 0 and np.isfinite(u2).all()
    # the same warm-started problem on the oracle: U_last = U guess = previous U*, X guess = previous X*
    b2 = {k: (v[i:i + 1] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == B else v) for k, v in b.items()}
    b2["u_last"] = ul[None]; b2["x_guess"] = xg[None]
    r2 = _oracle_solve(b2)
    assert r2["status"][0] == 0
    assert abs(c1.cost - r2["cost"][0]) <= 1e-5 * max(1.0, abs(r2["cost"][0])) and np.abs(u2 - r2["U"][0, 0]).max() < 1e-4
    # the model is refused where it cannot run: planes, or the literal-reference NLP mode
    from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
    from mobile_manipulator_mpc_b200._lib import MmpcError
    with pytest.raises(MmpcError):
        BatchSolver(N=10, n_obs=3, n_pl=2, mode=_abi.MODE_CLEAN, model=_abi.MODEL_POSEREF)
    with pytest.raises(MmpcError):
        BatchSolver(N=10, n_obs=3, n_pl=0, mode=_abi.MODE_REFERENCE, model=_abi.MODEL_POSEREF)
