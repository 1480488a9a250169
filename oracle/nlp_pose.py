"""CPU oracle -- the NLP of the reference's pose-reference whole-body controller (controllers/mpc_wholebody.py:49-128: the
tracking cost is on the END-POINT POSE (x, y, z, psi) through forward_tranformation, not on the state), restated densely in
NumPy, one instance.

TEST INFRASTRUCTURE ONLY.  Function values pinned to the reference's own reset() by tests/golden/ref_rows_pose.npz
(tests/golden/make_ref_rows.py::make_pose_case); solutions by the complex-step KKT check of tests/test_mpc_pose.py and the
oracle's interior-point solver (no IPOPT in this image).

Decision vector  w = [X[1..N] (9N), U (5N), s (N+1)]   (X[0] == X_init substituted, :108).  Rows: dynamics (:76), boxes on
u (:91), x (:92, :109) and u - u_last (:93), ground circles against the slack of their stage (:96-97, :112-113).  No
self-collision rows, no planes ("onstacles 3D: TODO", :100, :116).
"""
import numpy as np

from . import model as M

PI, INF = np.pi, np.inf


class NLPPose:
    def __init__(self, N, dt, x_init, x_ref, u_ref, u_last, circles, Qd=None, Pd=None, Rd=None, Wd=None, S=1e5,
                 ulim=None, xlim=None, dulim=None, base_radius=0.4):
        self.N, self.dt = N, dt
        self.Qd = 5 * np.ones(4) if Qd is None else np.asarray(Qd, float)                      # :11  x, y, z, psi of the end point
        self.Pd = 50 * np.ones(4) if Pd is None else np.asarray(Pd, float)                     # :12
        self.Rd = np.array([0.1, 0.1, 0, 0, 0]) if Rd is None else np.asarray(Rd, float)       # :13
        self.Wd = np.array([0, 0, 0.1, 0.1, 0.1]) if Wd is None else np.asarray(Wd, float)     # :15
        self.S = float(S)                                                                       # :14
        self.ulim = np.array([[-2, -PI, -1, -1, -1], [2, PI, 1, 1, 1]]) if ulim is None else np.asarray(ulim, float)   # :16
        self.xlim = (np.array([[-100, -100, -INF, -2, -2, -PI, -PI / 2, -PI * 3 / 4, 0],
                               [100, 100, INF, 2, 2, PI, PI / 2, 0, PI]]) if xlim is None else np.asarray(xlim, float))  # :17-20
        self.dulim = (np.array([[-INF, -INF, -0.5, -0.5, -0.5], [INF, INF, 0.5, 0.5, 0.5]])
                      if dulim is None else np.asarray(dulim, float))                           # :21
        self.x_init = np.clip(np.asarray(x_init, float), self.xlim[0], self.xlim[1])           # :133-134
        self.x_ref = np.asarray(x_ref, float)[:, :4]                                            # [N+1, 4]  :66
        self.u_ref, self.u_last = np.asarray(u_ref, float), np.asarray(u_last, float)
        self.circles = np.asarray(circles, float).reshape(-1, 3)
        self.base_radius = base_radius
        self.nw = 9 * N + 5 * N + N + 1

    def unpack(self, w):
        N = self.N
        X = np.concatenate([self.x_init[None].astype(w.dtype), w[:9 * N].reshape(N, 9)])
        return X, w[9 * N:14 * N].reshape(N, 5), w[14 * N:]

    def pack(self, X, U, s):
        return np.concatenate([np.asarray(X)[1:].ravel(), np.asarray(U).ravel(), np.ravel(s)])

    def initial_guess(self, x_guess=None, u_guess=None):
        """:138-146: X <- x_guess (tile(x_init) on the first call), U <- u_latest (zeros), s <- 0."""
        N = self.N
        X = np.tile(self.x_init, (N + 1, 1)) if x_guess is None else np.asarray(x_guess, float)
        return self.pack(X, self.u_last if u_guess is None else u_guess, np.zeros(N + 1))

    def pose_error(self, X):
        """:79-80, :104-105: forward_tranformation(x)[0] - X_ref, rows of (x, y, z, psi); psi plainly (no angleDiff here)."""
        pe, _, _ = M.forward_transformation(X)
        return np.stack([pe[0], pe[1], pe[2] + 0 * pe[0], pe[3]], axis=-1) - self.x_ref

    def cost(self, w):
        X, U, s = self.unpack(w)
        N = self.N
        e = self.pose_error(X)
        J = np.sum(self.Qd * e[:N] ** 2) + np.sum(self.Pd * e[N] ** 2)                                     # :86, :107
        J = J + np.sum(self.Rd * (U - self.u_ref) ** 2) + np.sum(self.Wd * (U - self.u_last) ** 2)         # :87-88
        return J + self.S * np.sum(s ** 2)                                                                 # :98, :114

    def eq(self, w):
        X, U, _ = self.unpack(w)
        return (M.f_kinematics(X[:-1], U, self.dt) - X[1:]).ravel()                                       # :76

    def circle_rows(self, X, s):
        rows = []
        for k in range(self.N + 1):
            for (ox, oy, r) in self.circles:
                rows.append((r + self.base_radius) - np.sqrt((X[k, 0] - ox) ** 2 + (X[k, 1] - oy) ** 2) + 0.0 - s[k])   # :42-46
        return np.array(rows, dtype=X.dtype) if rows else np.zeros(0, dtype=X.dtype)

    def ineq(self, w, with_boxes=True):
        """all rows as g(w) <= 0: circles (stage-major), then the finite boxes on x[1..N], u, u - u_last."""
        X, U, s = self.unpack(w)
        g = [self.circle_rows(X, s)]
        if with_boxes:
            fx = np.isfinite(self.xlim); fd = np.isfinite(self.dulim)
            g += [(self.xlim[0] - X[1:])[:, fx[0]].ravel(), (X[1:] - self.xlim[1])[:, fx[1]].ravel(),
                  (self.ulim[0] - U).ravel(), (U - self.ulim[1]).ravel(),
                  (self.dulim[0] - (U - self.u_last))[:, fd[0]].ravel(), ((U - self.u_last) - self.dulim[1])[:, fd[1]].ravel()]
        return np.concatenate(g)

    def jac(self, fun, w):
        f0 = np.atleast_1d(fun(w))
        J = np.empty((f0.size, w.size))
        for i in range(w.size):
            wc = w.astype(complex); wc[i] += 1e-30j
            J[:, i] = np.atleast_1d(fun(wc)).imag / 1e-30
        return J

    def violation(self, w):
        return max(np.abs(self.eq(w)).max(), self.ineq(w).max(), 0.0)

    def kkt_residual(self, w, tol_active=1e-6):
        """Stationarity of a candidate optimum by non-negative least squares over the multipliers of the active rows (complex-step
        Jacobians): min || grad f + Jeq^T lam + Jact^T z ||, z >= 0."""
        from scipy.optimize import nnls
        g = self.jac(self.cost, w)[0]
        Je = self.jac(self.eq, w)
        gi = self.ineq(w)
        act = gi >= -tol_active
        Ji = self.jac(self.ineq, w)[act]
        A = np.concatenate([Je.T, -Je.T, Ji.T], axis=1)
        sol, res = nnls(A, -g)
        return res / max(1.0, np.abs(g).max())


def from_batch(batch, b=0, **kw):
    """The instance b of a batch in the layout of scenarios.make_pose_batch."""
    for k in ("Qd", "Pd", "Rd", "Wd", "S"):
        if k in batch and k not in kw:
            kw[k] = np.asarray(batch[k], float)[:4] if k in ("Qd", "Pd") else batch[k]
    kw.setdefault("xlim", None)
    return NLPPose(batch["N"], batch["dt"], batch["x_init"][b], batch["x_ref"][b], batch["u_ref"][b], batch["u_last"][b],
                   batch["circles"][b] if batch.get("circles") is not None else np.zeros((0, 3)), **kw)
