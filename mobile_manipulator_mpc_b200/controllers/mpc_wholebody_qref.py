"""Drop-in ``MPCWholeBody`` -- same constructor, attributes and ``reset()/solve()`` contract as
the reference's controllers/mpc_wholebody_qref.py:6-331, with ``opti.solve()`` (:315) replaced by
the batched sm_100a interior-point kernel behind include/mmpc.h.

What the caller (interface_wholebody_qref.py) touches and what it gets here:
  solve(x_init, traj_ref, u_ref) -> (5,) float64, U*[0]   :287-331  (x_init[6:] clipped IN PLACE, :290)
  setWeight(Q=, R=, P=, S=, W=)                            :119-139
  angleDiff(a, b) -> float                                 :92-117
  .N .dt .f_dynamics .robot_model .obstacle_list .x_guess .u_latest .ulim .xlim .dulim .base_radius
  .opti.subject_to(.X[N, :2] == .X_ref[N, :2])             interface_wholebody_qref.py:167 (shim)
Extra (keyword-only): batch, device, mode ("reference" = bug-for-bug NLP, the default; "clean" = stage-separable),
and ``solve_batch`` for B instances at once.
"""
import numpy as np

from .. import _abi
from ..batch_solver import BatchSolver

PI = np.pi


class _SymSlice:
    def __init__(self, name, key):
        self.name, self.key = name, key

    def __eq__(self, other):
        return _EqExpr(self, other)

    __hash__ = None


class _EqExpr:
    def __init__(self, lhs, rhs):
        self.lhs, self.rhs = lhs, rhs


class _SymMatrix:
    """Stand-in for the Opti variable/parameter objects the Interface indexes (X, X_ref)."""

    def __init__(self, name, shape):
        self.name, self.shape = name, shape

    def __getitem__(self, key):
        return _SymSlice(self.name, key)


class _OptiShim:
    """Recognises exactly the one NLP mutation the reference's caller performs:
    ``opti.subject_to(X[N, :2] == X_ref[N, :2])`` (interface_wholebody_qref.py:167)."""

    def __init__(self, ctrl):
        self._c = ctrl

    def subject_to(self, expr):
        c = self._c
        ok = (isinstance(expr, _EqExpr) and isinstance(expr.lhs, _SymSlice) and isinstance(expr.rhs, _SymSlice)
              and expr.lhs.name == "X" and expr.rhs.name == "X_ref"
              and expr.lhs.key == (c.N, slice(None, 2, None)) and expr.rhs.key == (c.N, slice(None, 2, None)))
        if not ok:
            raise NotImplementedError("only opti.subject_to(X[N, :2] == X_ref[N, :2]) is supported")
        c.terminal_xy_eq = True


class MPCWholeBody:
    def __init__(self,
                 robot,
                 obstacle_list,
                 obstacle_manipulation_list,
                 N=10,
                 Q=5 * np.diag([5, 5, 0, 0, 0, 1, 1, 1, 1]),    # x y psi dx dy dpsi q1 q2 q3
                 P=5 * np.diag([5, 5, 0, 0, 0, 1, 1, 1, 1]),
                 R=np.diag([0.1, 0.1, 0.0, 0.0, 0.0]),          # dV, dw, dq1, dq2, dq3
                 S=np.diag([1e5]),                              # slack variable s, cost += S*s**2
                 W=np.diag([0, 0, 1e-1, 1e-1, 1e-1]),           # ddV, ddw, ddq1, ddq2, ddq3
                 ulim=np.array([[-2, -PI, -1, -1, -1], [2, PI, 1, 1, 1]]),
                 xlim=np.array([[-100, -100, -np.inf, -2, -2, -PI, -PI / 2, -PI, 0],
                                [100, 100, np.inf, 2, 2, PI, PI / 2, 0, 3 * PI / 2]]),
                 dulim=np.array([[-np.inf, -np.inf, -0.5, -0.5, -0.5], [np.inf, np.inf, 0.5, 0.5, 0.5]]),
                 *, batch=1, device=0, mode="reference", verbose=True):
        self.N = N
        self.Q_value, self.R_value, self.P_value, self.S_value, self.W_value = Q, R, P, S, W
        self.dt = robot.dt
        self.dulim, self.ulim, self.xlim = np.asarray(dulim, float), np.asarray(ulim, float), np.asarray(xlim, float)
        self.f_dynamics = robot.f_kinematics
        self.robot_model = robot
        self.base_radius = robot.base.base_radius()
        self.obstacle_list = obstacle_list
        self.obstacle_manipulation_list = obstacle_manipulation_list
        self.endpoint_self_collision_radius = 0.05
        self.obstacle_expand_dist = 0.03
        self.batch, self.device, self.verbose = int(batch), int(device), verbose
        if mode not in ("clean", "reference"):
            raise ValueError("mode must be 'clean' or 'reference'")
        self.mode = mode
        self._solver = None
        self.reset()

    # -- reference helpers ---------------------------------------------------------------------
    def obsAvoid(self, obstacle_list, x):
        """(r + base_radius) - distance for every ground circle (:49-54); numeric."""
        x = np.asarray(x, dtype=float).reshape(-1)
        return [(o.radius + self.base_radius) - np.sqrt((x[0] - o.x) ** 2 + (x[1] - o.y) ** 2) + 0.0
                for o in obstacle_list]

    def angleDiff(self, a, b):
        """a - b wrapped to the nearest representative (:92-117), plain floats."""
        a = np.fmod(a + PI, 2 * PI) - PI
        b = np.fmod(b + PI, 2 * PI) - PI
        d = a - b
        if a * b >= 0:
            return float(d)
        if a > b:
            return float(d if d <= PI else d - 2 * PI)
        return float(d if d > -PI else d + 2 * PI)

    def setWeight(self, Q=None, R=None, P=None, S=None, W=None):
        """:119-139."""
        if Q is not None:
            self.Q_value = Q
        if R is not None:
            self.R_value = R
        if P is not None:
            self.P_value = P
        if S is not None:
            self.S_value = S
        if W is not None:
            self.W_value = W
        from ..batch_solver import _diag
        self.weights = dict(Qd=_diag(self.Q_value, 9, "Q"), Pd=_diag(self.P_value, 9, "P"), Rd=_diag(self.R_value, 5, "R"),
                            Wd=_diag(self.W_value, 5, "W"), S=float(np.asarray(self.S_value).reshape(-1)[0]))
        if self._solver is not None:
            self._solver.set_weights(Q=self.Q_value, R=self.R_value, P=self.P_value, S=self.S_value, W=self.W_value)

    # -- NLP definition --------------------------------------------------------------------------
    def _planes_array(self):
        pl = [np.hstack([np.asarray(p, float).reshape(3), np.asarray(n, float).reshape(3)])
              for p, n in self.obstacle_manipulation_list]
        return np.array(pl, dtype=float).reshape(len(pl), 6)

    def _circles_array(self):
        return np.array([[o.x, o.y, o.radius] for o in self.obstacle_list], dtype=float).reshape(len(self.obstacle_list), 3)

    def reset(self):
        """:142-285 -- fixes the NLP shape and (re)creates the device solver; clears the warm start."""
        n_pl = len(self.obstacle_manipulation_list)
        if n_pl > _abi.MAX_PLANES:
            raise NotImplementedError(f"at most {_abi.MAX_PLANES} planes")
        cfg = _abi.default_config(N=self.N, dt=self.dt, n_obs=len(self.obstacle_list), n_pl=n_pl,
                                  mode=_abi.MODE_CLEAN if self.mode == "clean" else _abi.MODE_REFERENCE)
        _abi.set_limits(cfg, ulim=self.ulim, xlim=self.xlim, dulim=self.dulim)
        cfg.base_radius = self.base_radius
        cfg.self_collision_radius = self.endpoint_self_collision_radius
        cfg.obstacle_expand_dist = self.obstacle_expand_dist
        if self._solver is not None:
            self._solver.close()
        self._solver = None          # the device handle is created by the first solve (needs a B200)
        self._cfg = cfg
        self.opti = _OptiShim(self)
        self.X = _SymMatrix("X", (self.N + 1, 9))
        self.U = _SymMatrix("U", (self.N, 5))
        self.s = _SymMatrix("s", (self.N + 1, 1))
        self.X_ref = _SymMatrix("X_ref", (self.N + 1, 9))
        self.U_ref = _SymMatrix("U_ref", (self.N, 5))
        self.terminal_xy_eq = False
        self.x_guess = None
        self.u_latest = None
        self.cost = None
        self.last_info = None
        self.setWeight()

    # -- the hot path ------------------------------------------------------------------------------
    def solve(self, x_init, traj_ref, u_ref):
        """:287-331.  Returns U*[0] (5,) float64; raises RuntimeError when the solve did not
        converge (the reference dies at :329 in that case)."""
        # :290-292
        x_init[6:] = np.maximum(np.minimum(x_init[6:], self.xlim[1, 6:]), self.xlim[0, 6:]).squeeze()
        x_init = np.maximum(np.minimum(x_init, self.xlim[1]), self.xlim[0]).squeeze()
        assert x_init[7] <= 0 and x_init[8] >= 0
        if self.x_guess is None:
            self.x_guess = np.ones((self.N + 1, 9)) * x_init
        if self.u_latest is None:
            self.u_latest = np.zeros((self.N, 5))
        out = self.solve_batch(x_init[None], np.asarray(traj_ref, float)[None], np.asarray(u_ref, float)[None],
                               u_last=self.u_latest[None])
        st = int(out["status"][0])
        if st not in (_abi.STATUS_CONVERGED, _abi.STATUS_ACCEPTABLE):
            raise RuntimeError(f"MPC solve failed: {_abi.STATUS_NAMES[st]} after {int(out['iters'][0])} iterations "
                               f"(KKT error {float(out['kkt'][0]):.3e})")
        self.cost = float(out["cost"][0])
        if self.verbose:
            print("cost: ", self.cost)   # :317
        self.x_guess = out["X"][0]       # :329
        self.u_latest = out["U"][0]      # :330
        return self.u_latest[0, :]

    def _backend_solve(self, batch, B):
        """The one call into the C ABI (mmpc_solve_host); the product has no other backend."""
        if self._solver is None:
            self._solver = BatchSolver(cfg=self._cfg, B_max=max(self.batch, B), device=self.device)
            self._solver.set_weights(Q=self.Q_value, R=self.R_value, P=self.P_value, S=self.S_value, W=self.W_value)
        return self._solver.solve_host(batch)

    def solve_batch(self, x_init, traj_ref, u_ref, u_last=None, u_guess=None, circles=None, planes=None,
                    n_pl_inst=None):
        """B instances at once (host arrays).  circles/planes default to the constructor's lists."""
        x_init = np.asarray(x_init, float)
        B = x_init.shape[0]
        N = self.N
        batch = dict(x_init=x_init, x_ref=traj_ref, u_ref=u_ref,
                     u_last=np.zeros((B, N, 5)) if u_last is None else u_last, u_guess=u_guess)
        c = self._cfg
        if c.n_obs:
            batch["circles"] = np.broadcast_to(self._circles_array(), (B, c.n_obs, 3)) if circles is None else circles
        if c.n_pl:
            batch["planes"] = np.broadcast_to(self._planes_array(), (B, c.n_pl, 6)) if planes is None else planes
        batch["n_pl_inst"] = n_pl_inst
        if self.terminal_xy_eq:   # opti.subject_to(X[N, :2] == X_ref[N, :2])  interface_wholebody_qref.py:167
            batch["flags"] = np.ones(B, np.uint8)
        out = self._backend_solve(batch, B)
        self.last_info = {k: out[k] for k in ("status", "iters", "kkt", "cost")}
        return out
