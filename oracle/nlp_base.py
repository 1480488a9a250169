"""CPU oracle -- the NLP of the reference's base-only controller MPCBase (controllers/mpc_base.py:114-189) restated densely
in NumPy, one instance, native 6-state / 2-control dimensions (independent of the 9 / 5 embedding the solvers use).

TEST INFRASTRUCTURE ONLY.  Function values pinned to the reference's own reset() by tests/golden/ref_rows_base.npz
(tests/golden/make_ref_rows.py); solutions by SciPy SLSQP on this restatement (tests/golden/make_golden_base.py).

Decision vector  w = [X[1..N] (6N), U (2N), s (N+1)]   (X[0] == X_init substituted, :153).
"""
import numpy as np

from . import model as M

PI = np.pi


def angle_diff(a, b):
    """mpc_base.py:59-84, vectorised over arrays (real or complex-step)."""
    a = np.asarray(a); b = np.asarray(b)
    aw = np.fmod(a.real + PI, 2 * PI) - PI + 1j * a.imag if np.iscomplexobj(a) else np.fmod(a + PI, 2 * PI) - PI
    bw = np.fmod(b + PI, 2 * PI) - PI
    d = aw - bw
    ar, dr = np.real(aw), np.real(d)
    same = ar * bw >= 0
    out = np.where(same, d, np.where(ar > bw, np.where(dr <= PI, d, d - 2 * PI), np.where(dr > -PI, d, d + 2 * PI)))
    return out


class NLPBase:
    def __init__(self, N, dt, x_init, x_ref, u_ref, circles, Qd=None, Pd=None, Rd=None, M_=1e5, ulim=None, xlim=None, base_radius=0.4):
        self.N, self.dt = N, dt
        self.Qd = np.array([5., 5., 0, 0, 0, 1.]) if Qd is None else np.asarray(Qd, float)     # :11
        self.Pd = self.Qd.copy() if Pd is None else np.asarray(Pd, float)                       # :12
        self.Rd = np.array([1., 1.]) if Rd is None else np.asarray(Rd, float)                   # :13
        self.M = float(M_)                                                                      # :14
        self.ulim = np.array([[-2, -PI], [2, PI]]) if ulim is None else np.asarray(ulim, float)                  # :15
        self.xlim = np.array([[-100, -100, -2, -2, -PI], [100, 100, 2, 2, PI]]) if xlim is None else np.asarray(xlim, float)   # :16
        self.x_init = np.asarray(x_init, float)
        self.x_ref, self.u_ref = np.asarray(x_ref, float), np.asarray(u_ref, float)
        self.circles = np.asarray(circles, float).reshape(-1, 3)
        self.base_radius = base_radius
        self.nw = 6 * N + 2 * N + N + 1

    def unpack(self, w):
        N = self.N
        X = np.concatenate([self.x_init[None].astype(w.dtype), w[:6 * N].reshape(N, 6)])
        return X, w[6 * N:8 * N].reshape(N, 2), w[8 * N:]

    def pack(self, X, U, s):
        return np.concatenate([X[1:].ravel(), U.ravel(), np.ravel(s)])

    def initial_guess(self, x_guess=None, u_guess=None):
        """:193-201: X <- X_guess (tile(x_init) on the first call), U <- U_guess (zeros), s <- 0."""
        N = self.N
        X = np.tile(self.x_init, (N + 1, 1)) if x_guess is None else np.asarray(x_guess, float)
        return self.pack(X, np.zeros((N, 2)) if u_guess is None else u_guess, np.zeros(N + 1))

    def state_error(self, X):
        """:129-133, :146-150: plain differences, the yaw through angleDiff."""
        e = X - self.x_ref
        e = e.astype(X.dtype)
        e[:, 2] = angle_diff(X[:, 2], self.x_ref[:, 2])
        return e

    def cost(self, w):
        X, U, s = self.unpack(w)
        N = self.N
        e = self.state_error(X)
        J = np.sum(self.Qd * e[:N] ** 2) + np.sum(self.Pd * e[N] ** 2) + np.sum(self.Rd * (U - self.u_ref) ** 2)    # :135-136, :152
        return J + self.M * np.sum(s ** 2)                                                                             # :145, :160

    def dynamics(self, x, u):
        """robot_models/base.py:17-31."""
        dt = self.dt
        return np.stack([x[..., 0] + dt * x[..., 3], x[..., 1] + dt * x[..., 4], x[..., 2] + dt * x[..., 5],
                         x[..., 3] + dt * (u[..., 0] * np.cos(x[..., 2]) - x[..., 4] * x[..., 5]),
                         x[..., 4] + dt * (u[..., 0] * np.sin(x[..., 2]) + x[..., 3] * x[..., 5]),
                         x[..., 5] + dt * u[..., 1]], axis=-1)

    def eq(self, w):
        X, U, _ = self.unpack(w)
        return (self.dynamics(X[:-1], U) - X[1:]).ravel()                                                              # :128

    def ineq(self, w, with_boxes=True):
        """all rows as g(w) <= 0: circles (k-major) :142-143, :156-157; then the boxes :139-141, :154-155."""
        X, U, s = self.unpack(w)
        rows = []
        for k in range(self.N + 1):
            for (ox, oy, r) in self.circles:
                rows.append((r + self.base_radius) - np.sqrt((X[k, 0] - ox) ** 2 + (X[k, 1] - oy) ** 2) + 0.0 - s[k])
        g = [np.array(rows, dtype=w.dtype)] if rows else [np.zeros(0, dtype=w.dtype)]
        if with_boxes:
            xs = X[1:][:, [0, 1, 3, 4, 5]]
            g += [(self.xlim[0] - xs).ravel(), (xs - self.xlim[1]).ravel(), (self.ulim[0] - U).ravel(), (U - self.ulim[1]).ravel()]
        return np.concatenate(g)

    def jac(self, fun, w):
        f0 = fun(w)
        J = np.empty((f0.size, w.size))
        for i in range(w.size):
            wc = w.astype(complex); wc[i] += 1e-30j
            J[:, i] = fun(wc).imag / 1e-30
        return J

    def violation(self, w):
        return max(np.abs(self.eq(w)).max(), self.ineq(w).max(), 0.0)
