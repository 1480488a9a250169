"""Drop-in ``MPCBase`` -- the base-only controller of the reference (controllers/mpc_base.py:6-229):
same constructor, ``reset()/solve()/setWeight()/angleDiff()/obsAvoid()`` and attributes, with ``opti.solve()`` (:209)
replaced by the batched sm_100a interior-point kernels behind include/mmpc.h (``MmpcConfig.model = MMPC_MODEL_BASE``).

The NLP (:114-189): states x y psi dx dy dpsi, controls dV dw, one free slack per stage;
    min  sum_k e_k Q e_k' + (u_k - ur_k) R (u_k - ur_k)' + M s_k^2  +  e_N P e_N' + M s_N^2,
         e_k = [x - xr, y - yr, angleDiff(psi, psir), dx - dxr, dy - dyr, dpsi - dpsir]                    (:129-141, :146-152)
    s.t. X[k+1] = Base.f_kinematics(X[k], U[k])   (:128)      X[0] = X_init  (:153)
         ulim on U[k] (:139), xlim on (x, y) and on (dx, dy, dpsi) of every X[k] (:140-141, :154-155), psi free
         (r_i + base_radius) - ||(x, y) - o_i|| <= s_k for every ground circle (:142-143, :156-157)
The library solves it as the whole-body NLP with the arm taken out (see include/mmpc.h, MMPC_MODEL_BASE): the arrays keep the
9 / 5 layout, the arm entries are zero, unbounded, and decoupled from the base.  Warm start as in :196-201: X and U start
from the previous solution (zeros / tile(x_init) on the first call)."""
import numpy as np

from .. import _abi
from ..batch_solver import BatchSolver, _diag

PI = np.pi


class MPCBase:
    def __init__(self,
                 robot,
                 obstacle_list,
                 N=10,
                 Q=np.diag([5., 5., 0.0, 0, 0, 1.]),    # x y psi dx dy dpsi
                 P=np.diag([5., 5., 0.0, 0, 0, 1.]),
                 R=np.diag([1., 1.]),
                 M=np.diag([1e5]),
                 ulim=np.array([[-2, -PI], [2, PI]]),                                   # dv, dw
                 xlim=np.array([[-100, -100, -2, -2, -PI], [100, 100, 2, 2, PI]]),      # x, y, _, dx, dy, dpsi
                 *, batch=1, device=0, verbose=False):
        self.Q_value, self.R_value, self.P_value, self.M_value = Q, R, P, M
        self.dt = robot.dt
        self.N = N
        self.ulim, self.xlim = np.asarray(ulim, float), np.asarray(xlim, float)
        self.f_dynamics = robot.f_kinematics      # member function
        self.base_radius = robot.base_radius      # member function (the reference keeps the bound method, :28)
        self.obstacle_list = obstacle_list
        self.batch, self.device, self.verbose = int(batch), int(device), verbose
        self._solver = None
        self.reset()

    # -- reference helpers ---------------------------------------------------------------------
    def obsAvoid(self, obstacle_list, x):
        """(r + base_radius) - distance for every ground circle (:50-55); numeric."""
        x = np.asarray(x, dtype=float).reshape(-1)
        return [(o.radius + self.base_radius()) - np.sqrt((x[0] - o.x) ** 2 + (x[1] - o.y) ** 2) + 0.0 for o in obstacle_list]

    def angleDiff(self, a, b):
        """a - b wrapped to the nearest representative (:59-84), plain floats."""
        a = np.fmod(a + PI, 2 * PI) - PI
        b = np.fmod(b + PI, 2 * PI) - PI
        d = a - b
        if a * b >= 0:
            return float(d)
        if a > b:
            return float(d if d <= PI else d - 2 * PI)
        return float(d if d > -PI else d + 2 * PI)

    def setWeight(self, Q=None, R=None, P=None, M=None):
        """:86-112."""
        if Q is not None:
            self.Q_value = Q
        if R is not None:
            self.R_value = R
        if P is not None:
            self.P_value = P
        if M is not None:
            self.M_value = M
        self.weights = dict(Qd=np.concatenate([_diag(self.Q_value, 6, "Q"), np.zeros(3)]),
                            Pd=np.concatenate([_diag(self.P_value, 6, "P"), np.zeros(3)]),
                            Rd=np.concatenate([_diag(self.R_value, 2, "R"), np.ones(3)]),   # the absent arm: unit weight, stays put
                            Wd=np.zeros(5), S=float(np.asarray(self.M_value).reshape(-1)[0]))
        if self._solver is not None:
            self._push_weights()

    def _push_weights(self):
        w = self.weights
        self._solver.set_weights(Q=w["Qd"], P=w["Pd"], R=w["Rd"], W=w["Wd"], S=w["S"])

    def _circles_array(self):
        return np.array([[o.x, o.y, o.radius] for o in self.obstacle_list], dtype=float).reshape(len(self.obstacle_list), 3)

    def reset(self):
        """:114-189 -- fixes the NLP shape; clears the warm start."""
        inf = np.inf
        cfg = _abi.default_config(N=self.N, dt=self.dt, n_obs=len(self.obstacle_list), n_pl=0, mode=_abi.MODE_CLEAN)
        cfg.model = _abi.MODEL_BASE
        xl, ul = self.xlim, self.ulim
        _abi.set_limits(cfg,
                        ulim=np.array([[ul[0, 0], ul[0, 1], -inf, -inf, -inf], [ul[1, 0], ul[1, 1], inf, inf, inf]]),
                        xlim=np.array([[xl[0, 0], xl[0, 1], -inf, xl[0, 2], xl[0, 3], xl[0, 4], -inf, -inf, -inf],
                                       [xl[1, 0], xl[1, 1], inf, xl[1, 2], xl[1, 3], xl[1, 4], inf, inf, inf]]),
                        dulim=np.array([[-inf] * 5, [inf] * 5]))
        cfg.base_radius = self.base_radius()
        if self._solver is not None:
            self._solver.close()
        self._solver = None
        self._cfg = cfg
        self.X_guess = None
        self.U_guess = None
        self.last_info = None
        self.setWeight()

    # -- the hot path ------------------------------------------------------------------------------
    def solve(self, x_init, traj_ref, u_ref):
        """:191-229.  Returns U*[0] (2,) float64; raises RuntimeError when the solve did not converge (the reference dies at :225)."""
        x_init = np.asarray(x_init, float).reshape(6)
        if self.X_guess is None:
            self.X_guess = np.ones((self.N + 1, 6)) * x_init
        if self.U_guess is None:
            self.U_guess = np.zeros((self.N, 2))
        out = self.solve_batch(x_init[None], np.asarray(traj_ref, float)[None], np.asarray(u_ref, float)[None],
                               x_guess=self.X_guess[None], u_guess=self.U_guess[None])
        st = int(out["status"][0])
        if st not in (_abi.STATUS_CONVERGED, _abi.STATUS_ACCEPTABLE):
            raise RuntimeError(f"MPC solve failed: {_abi.STATUS_NAMES[st]} after {int(out['iters'][0])} iterations "
                               f"(KKT error {float(out['kkt'][0]):.3e})")
        self.cost = float(out["cost"][0])
        self.X_guess = out["X"][0]       # :224
        self.U_guess = out["U"][0]       # :225
        return self.U_guess[0, :]

    def _backend_solve(self, batch, B):
        """The one call into the C ABI (mmpc_solve_host); the product has no other backend."""
        if self._solver is None:
            self._solver = BatchSolver(cfg=self._cfg, B_max=max(self.batch, B), device=self.device)
            self._push_weights()
        return self._solver.solve_host(batch)

    def solve_batch(self, x_init, traj_ref, u_ref, x_guess=None, u_guess=None, circles=None):
        """B instances at once: x_init [B,6], traj_ref [B,N+1,6], u_ref [B,N,2]; optional warm starts x_guess [B,N+1,6],
        u_guess [B,N,2].  Returns dict with U [B,N,2], X [B,N+1,6], s, cost, kkt, iters, status."""
        x_init = np.asarray(x_init, float)
        B, N = x_init.shape[0], self.N
        pad = lambda a, n: None if a is None else np.concatenate([np.asarray(a, float), np.zeros(a.shape[:-1] + (n,))], axis=-1)
        batch = dict(x_init=pad(x_init, 3), x_ref=pad(np.asarray(traj_ref, float), 3), u_ref=pad(np.asarray(u_ref, float), 3),
                     u_last=np.zeros((B, N, 5)), u_guess=pad(u_guess, 3), x_guess=pad(x_guess, 3))
        c = self._cfg
        if c.n_obs:
            batch["circles"] = np.broadcast_to(self._circles_array(), (B, c.n_obs, 3)) if circles is None else circles
        out9 = self._backend_solve(batch, B)
        out = dict(out9)
        out["U"] = np.ascontiguousarray(out9["U"][:, :, :2])
        out["X"] = np.ascontiguousarray(out9["X"][:, :, :6])
        self.last_info = {k: out[k] for k in ("status", "iters", "kkt", "cost")}
        return out
