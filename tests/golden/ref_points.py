"""Random evaluation points of the reference-row goldens (shared by make_ref_rows.py and the tests that read
tests/golden/ref_rows_*.npz: the points are regenerated from the stored seed, the files keep their checksums)."""
import numpy as np

PI = np.pi


def random_points(rng, M, N, npl):
    lo = np.array([-1, -1, -PI, -2, -2, -PI, -PI / 2, -PI, 0.0])
    hi = np.array([6, 6, PI, 2, 2, PI, PI / 2, 0, 3 * PI / 2])
    X = rng.uniform(lo, hi, size=(M, N + 1, 9))
    ul = np.array([2, PI, 1, 1, 1.0])
    U = rng.uniform(-ul, ul, size=(M, N, 5))
    s = rng.uniform(-0.1, 0.5, size=(M, N + 1, 1))
    free = rng.normal(size=(M, 6, npl))                       # the free `constr` decision variables (:156)
    U_last = rng.uniform(-ul, ul, size=(M, N, 5))
    X_ref = rng.uniform(lo, hi, size=(M, N + 1, 9))
    U_ref = rng.uniform(-ul, ul, size=(M, N, 5))
    return dict(X=X, U=U, s=s, free=free, U_last=U_last, X_init=X[:, 0:1, :].copy(), X_ref=X_ref, U_ref=U_ref)


def random_points_base(rng, M, N):
    """evaluation points of the MPCBase rows (6 states, 2 controls)"""
    lo = np.array([-1, -1, -2 * PI, -2, -2, -PI]); hi = np.array([6, 6, 2 * PI, 2, 2, PI])
    pts = dict(X=rng.uniform(lo, hi, size=(M, N + 1, 6)), U=rng.uniform([-2, -PI], [2, PI], size=(M, N, 2)),
               s=rng.uniform(-0.1, 0.5, size=(M, N + 1, 1)), X_ref=rng.uniform(lo, hi, size=(M, N + 1, 6)),
               U_ref=rng.uniform([-2, -PI], [2, PI], size=(M, N, 2)))
    pts["X_init"] = pts["X"][:, 0:1, :].copy()
    return pts


def input_checksums(pts):
    return np.array([float(np.sum(pts[k] * np.cos(np.arange(pts[k].size).reshape(pts[k].shape)))) for k in sorted(pts)])
