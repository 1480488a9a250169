"""CPU oracle -- the reference NLP restated densely in NumPy (one instance), independent of the
structured solver in mmpc_oracle.c.  Used to (a) check any candidate solution against the
restated constraints/cost, (b) produce golden solutions with SciPy (tests/golden/make_golden.py).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no CasADi/IPOPT in this image); follows
controllers/mpc_wholebody_qref.py:142-285 row by row (citations relative to /root/reference).

Decision vector  w = [X[1..N] (9N), U (5N), s (N+1)]   (X[0] == X_init is substituted, :244).
The free ``constr`` variables (:156) are dropped together with the k=0 rows they make vacuous
(SURVEY.md 8(a) row 8).  All functions accept complex ``w`` so that exact derivatives can be
taken by the complex-step method.
"""
import numpy as np

from . import model as M


class NLP:
    def __init__(self, N, dt, x_init, x_ref, u_ref, u_last, circles, planes, mode="reference",
                 Qd=None, Pd=None, Rd=None, Wd=None, S=1e5, ulim=None, xlim=None, dulim=None, terminal_xy_eq=False):
        pi, inf = np.pi, np.inf
        self.N, self.dt = N, dt
        self.Qd = np.array([25, 25, 0, 0, 0, 5, 5, 5, 5.0]) if Qd is None else np.asarray(Qd, float)   # :12
        self.Pd = self.Qd.copy() if Pd is None else np.asarray(Pd, float)                               # :13
        self.Rd = np.array([0.1, 0.1, 0, 0, 0]) if Rd is None else np.asarray(Rd, float)               # :14
        self.Wd = np.array([0, 0, 0.1, 0.1, 0.1]) if Wd is None else np.asarray(Wd, float)             # :16
        self.S = S                                                                                      # :15
        self.ulim = np.array([[-2, -pi, -1, -1, -1], [2, pi, 1, 1, 1]]) if ulim is None else ulim      # :17
        self.xlim = (np.array([[-100, -100, -inf, -2, -2, -pi, -pi / 2, -pi, 0],
                               [100, 100, inf, 2, 2, pi, pi / 2, 0, 3 * pi / 2]]) if xlim is None else xlim)  # :18-21
        self.dulim = (np.array([[-inf, -inf, -0.5, -0.5, -0.5], [inf, inf, 0.5, 0.5, 0.5]])
                      if dulim is None else dulim)                                                      # :22
        self.x_init = np.clip(np.asarray(x_init, float), self.xlim[0], self.xlim[1])                   # :290-291
        self.x_ref, self.u_ref, self.u_last = (np.asarray(a, float) for a in (x_ref, u_ref, u_last))
        circles = np.asarray(circles, float)
        self.circles = circles if circles.ndim == 3 else np.broadcast_to(circles, (N + 1,) + circles.shape)
        self.planes = [(np.asarray(p[:3], float), np.asarray(p[3:], float)) for p in planes]
        self.mode = mode
        self.terminal_xy_eq = bool(terminal_xy_eq)   # interface_wholebody_qref.py:167
        self.nw = 9 * N + 5 * N + N + 1

    # -- packing ---------------------------------------------------------------------------
    def unpack(self, w):
        N = self.N
        X = np.concatenate([self.x_init[None].astype(w.dtype), w[:9 * N].reshape(N, 9)])
        U = w[9 * N:14 * N].reshape(N, 5)
        s = w[14 * N:]
        return X, U, s

    def pack(self, X, U, s):
        return np.concatenate([X[1:].ravel(), U.ravel(), s.ravel()])

    def initial_guess(self, u_guess=None):
        """:302-304: X <- tile(x_init), U <- u_latest, s <- 0."""
        N = self.N
        return self.pack(np.tile(self.x_init, (N + 1, 1)), self.u_last if u_guess is None else u_guess,
                         np.zeros(N + 1))

    # -- objective :192-201,227,240-242,270 ---------------------------------------------------
    def cost(self, w):
        X, U, s = self.unpack(w)
        N = self.N
        ex = X - self.x_ref
        J = np.sum(self.Qd * ex[:N] ** 2) + np.sum(self.Pd * ex[N] ** 2)
        J = J + np.sum(self.Rd * (U - self.u_ref) ** 2) + np.sum(self.Wd * (U - self.u_last) ** 2)
        return J + self.S * np.sum(s ** 2)

    def cost_grad(self, w):
        X, U, s = self.unpack(w)
        N = self.N
        gx = 2 * self.Qd * (X - self.x_ref)
        gx[N] = 2 * self.Pd * (X[N] - self.x_ref[N])
        gu = 2 * self.Rd * (U - self.u_ref) + 2 * self.Wd * (U - self.u_last)
        return np.concatenate([gx[1:].ravel(), gu.ravel(), 2 * self.S * s])

    # -- equalities :180 ------------------------------------------------------------------------
    def eq(self, w):
        X, U, _ = self.unpack(w)
        e = (M.f_kinematics(X[:-1], U, self.dt) - X[1:]).ravel()
        if self.terminal_xy_eq:   # opti.subject_to(X[N, :2] == X_ref[N, :2])  interface_wholebody_qref.py:167
            e = np.concatenate([e, X[self.N, :2] - self.x_ref[self.N, :2]])
        return e

    # -- inequalities, all as g(w) <= 0 -----------------------------------------------------------
    def ineq(self, w, with_boxes=True):
        X, U, s = self.unpack(w)
        N = self.N
        rows = []
        # circles :208-209, :248-249
        for k in range(N + 1):
            for g in M.circle_rows(X[k], self.circles[k]):
                rows.append(g - s[k])
        # self collision :219-222 (stage), :261-265 (terminal, quirk 3: s[N-1])
        for k in range(N + 1):
            sk = s[k] if (k < N or self.mode == "clean") else s[N - 1]
            for g in M.self_collision_rows(X[k]):
                rows.append(g - sk)
        # planes :57-89
        if self.planes:
            npl = len(self.planes)
            c = [M.plane_margins(X[k], self.planes) for k in range(N + 1)]  # c[k][i][j]
            for k in range(N + 1):
                for i in range(6):
                    js = [npl - 1] if self.mode == "clean" else range(npl)
                    for j in js:
                        if j < npl - 1 and k == 0:
                            continue  # vacuous: the stale columns are free variables (quirk 2)
                        cols = [c[k][i][jj] if jj <= j else c[k - 1][i][jj] for jj in range(npl)]  # quirk 1
                        rows.append(-self._max(cols) - s[k])
        g = [np.stack(rows)]
        if with_boxes:
            g += self._boxes(X, U)
        return np.concatenate(g)

    @staticmethod
    def _max(cols):
        if len(cols) == 1:
            return cols[0]                                        # :82-83
        if len(cols) == 2:
            return cols[0] if cols[0].real > cols[1].real else cols[1]   # :84-85 if_else(c0 > c1, c0, c1)
        b = 0                                                      # :86-87 mmax
        for j in range(1, len(cols)):
            if cols[j].real > cols[b].real:
                b = j
        return cols[b]

    def _boxes(self, X, U):
        out = []
        N = self.N
        for lim, V in ((self.xlim, X[1:]), (self.ulim, U), (self.dulim, U - self.u_last)):   # :203-205, :245
            lo, hi = lim
            fl, fh = np.isfinite(lo), np.isfinite(hi)
            out.append((lo[fl] - V[:, fl]).ravel())
            out.append((V[:, fh] - hi[fh]).ravel())
        return out

    # -- derivatives by complex step --------------------------------------------------------------
    def jac(self, fun, w, chunk=64):
        n = w.size
        f0 = fun(w)
        J = np.empty((f0.size, n))
        h = 1e-30
        for i in range(n):
            wc = w.astype(complex)
            wc[i] += 1j * h
            J[:, i] = fun(wc).imag / h
        return J

    # -- report ---------------------------------------------------------------------------------
    def violation(self, w):
        """max violation of the restated NLP: equalities, g - s rows, boxes."""
        return max(np.abs(self.eq(w)).max(), self.ineq(w).max(), 0.0)

    def active_rows(self, w, thr=1e-6):
        """indices of obstacle rows (circles, self-collision, planes) with g - s >= -thr."""
        return np.nonzero(self.ineq(w, with_boxes=False) >= -thr)[0]


def from_batch(batch, b=0, mode="reference"):
    circ = batch["circles"][b]
    npl = int(batch["n_pl_inst"][b]) if batch.get("n_pl_inst") is not None else batch["n_pl"]
    kw = {}
    if "Qd" in batch:
        kw["Qd"] = batch["Qd"]; kw["Pd"] = batch.get("Pd", batch["Qd"])
    if batch.get("flags") is not None:
        kw["terminal_xy_eq"] = bool(int(batch["flags"][b]) & 1)
    return NLP(batch["N"], batch["dt"], batch["x_init"][b], batch["x_ref"][b], batch["u_ref"][b], batch["u_last"][b],
               circ, batch["planes"][b][:npl], mode=mode, **kw)
