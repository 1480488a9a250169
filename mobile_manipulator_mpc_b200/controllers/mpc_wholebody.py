"""Drop-in ``MPCWholeBody`` of the reference's POSE-REFERENCE controller (controllers/mpc_wholebody.py:6-175; SURVEY.md 8(f)
row 4; not to be confused with controllers/mpc_wholebody_qref.py, the controller of the demo): same constructor,
``reset()/solve()/obsAvoid()`` and attributes, with ``opti.solve()`` (:158) replaced by the sm_100a interior-point kernel behind
include/mmpc.h (``MmpcConfig.model = MMPC_MODEL_POSEREF``, resident kernel).

The NLP (:49-128): whole-body states and controls, one free slack per stage;
    min  sum_k e_k Q e_k' + (u_k - ur_k) R (u_k - ur_k)' + (u_k - ul_k) W (u_k - ul_k)' + S s_k^2  +  e_N P e_N' + S s_N^2,
         e_k = forward_tranformation(X[k])[0] - X_ref[k]      the END-POINT pose (x, y, z, psi), X_ref [N+1, 4]   (:79-80, :104-105)
    s.t. X[k+1] = f_kinematics(X[k], U[k]) (:76), X[0] = X_init (:108), ulim / xlim / dulim boxes (:91-93, :109),
         (r_i + base_radius) - ||(x, y) - o_i|| <= s_k for every ground circle (:96-97, :112-113); no 3-D obstacles (":100 TODO").
Warm start as in :138-146: X from the previous solution (tile(x_init) on the first call), U and U_last from the previous U*."""
import numpy as np

from .. import _abi
from ..batch_solver import BatchSolver, _diag

PI = np.pi


class MPCWholeBody:
    def __init__(self,
                 robot,
                 obstacle_list,
                 N=10,
                 Q=5 * np.diag([1, 1, 1, 1]),        # x, y, z, psi of the end point
                 P=50 * np.diag([1, 1, 1, 1]),
                 R=np.diag([0.1, 0.1, 0.0, 0.0, 0.0]),  # dV, dw, dq1, dq2, dq3
                 S=np.diag([1e5]),
                 W=np.diag([0, 0, 1e-1, 1e-1, 1e-1]),
                 ulim=np.array([[-2, -PI, -1, -1, -1], [2, PI, 1, 1, 1]]),
                 xlim=np.array([[-100, -100, -np.inf, -2, -2, -PI, -PI / 2, -PI * 3 / 4, 0],
                                [100, 100, np.inf, 2, 2, PI, PI / 2, 0, PI]]),
                 dulim=np.array([[-np.inf, -np.inf, -0.5, -0.5, -0.5], [np.inf, np.inf, 0.5, 0.5, 0.5]]),
                 *, batch=1, device=0, verbose=False):
        self.N = N
        self.Q, self.R, self.P, self.S, self.W = Q, R, P, S, W
        self.dt = robot.dt
        self.dulim, self.ulim, self.xlim = np.asarray(dulim, float), np.asarray(ulim, float), np.asarray(xlim, float)
        self.f_dynamics = robot.f_kinematics
        self.robot_model = robot
        self.base_radius = robot.base.base_radius()
        self.obstacle_list = obstacle_list
        self.batch, self.device, self.verbose = int(batch), int(device), verbose
        self._solver = None
        self.reset()

    def obsAvoid(self, obstacle_list, x):
        """:40-46, numeric."""
        x = np.asarray(x, dtype=float).reshape(-1)
        return [(o.radius + self.base_radius) - np.sqrt((x[0] - o.x) ** 2 + (x[1] - o.y) ** 2) + 0.0 for o in obstacle_list]

    def _circles_array(self):
        return np.array([[o.x, o.y, o.radius] for o in self.obstacle_list], dtype=float).reshape(len(self.obstacle_list), 3)

    def reset(self):
        """:49-128 -- fixes the NLP shape and weights; clears the warm start."""
        cfg = _abi.default_config(N=self.N, dt=self.dt, n_obs=len(self.obstacle_list), n_pl=0, mode=_abi.MODE_CLEAN)
        cfg.model = _abi.MODEL_POSEREF
        _abi.set_limits(cfg, ulim=self.ulim, xlim=self.xlim, dulim=self.dulim)
        cfg.base_radius = self.base_radius
        if self._solver is not None:
            self._solver.close()
        self._solver = None
        self._cfg = cfg
        self.weights = dict(Qd=np.concatenate([_diag(self.Q, 4, "Q"), np.zeros(5)]), Pd=np.concatenate([_diag(self.P, 4, "P"), np.zeros(5)]),
                            Rd=_diag(self.R, 5, "R"), Wd=_diag(self.W, 5, "W"), S=float(np.asarray(self.S).reshape(-1)[0]))
        self.x_guess = None
        self.u_latest = None
        self.cost = None
        self.last_info = None

    # -- the hot path ------------------------------------------------------------------------------
    def solve(self, x_init, traj_ref, u_ref):
        """:130-175.  traj_ref [N+1, 4]: end-point pose reference.  Returns U*[0] (5,) float64; raises RuntimeError when the solve
        did not converge (the reference prints its debug dump and dies at :172)."""
        x_init[6:] = np.maximum(np.minimum(x_init[6:], self.xlim[1, 6:]), self.xlim[0, 6:]).squeeze()   # :133
        x_init = np.maximum(np.minimum(x_init, self.xlim[1]), self.xlim[0]).squeeze()                     # :134
        assert x_init[7] <= 0 and x_init[8] >= 0                                                          # :135
        if self.x_guess is None:
            self.x_guess = np.ones((self.N + 1, 9)) * x_init
        if self.u_latest is None:
            self.u_latest = np.zeros((self.N, 5))
        out = self.solve_batch(x_init[None], np.asarray(traj_ref, float)[None], np.asarray(u_ref, float)[None],
                               u_last=self.u_latest[None], x_guess=self.x_guess[None])
        st = int(out["status"][0])
        if st not in (_abi.STATUS_CONVERGED, _abi.STATUS_ACCEPTABLE):
            raise RuntimeError(f"MPC solve failed: {_abi.STATUS_NAMES[st]} after {int(out['iters'][0])} iterations "
                               f"(KKT error {float(out['kkt'][0]):.3e})")
        self.cost = float(out["cost"][0])
        self.x_guess = out["X"][0]       # :173
        self.u_latest = out["U"][0]      # :174
        return self.u_latest[0, :]

    def _backend_solve(self, batch, B):
        """The one call into the C ABI (mmpc_solve_host); the product has no other backend."""
        if self._solver is None:
            self._solver = BatchSolver(cfg=self._cfg, B_max=max(self.batch, B), device=self.device)
            w = self.weights
            self._solver.set_weights(Q=w["Qd"], P=w["Pd"], R=w["Rd"], W=w["Wd"], S=w["S"])
        return self._solver.solve_host(batch)

    def solve_batch(self, x_init, traj_ref, u_ref, u_last=None, u_guess=None, x_guess=None, circles=None):
        """B instances at once: x_init [B,9], traj_ref [B,N+1,4] (end-point pose), u_ref [B,N,5]; optional u_last, warm starts
        u_guess [B,N,5] and x_guess [B,N+1,9].  Returns dict with U, X, s, cost, kkt, iters, status."""
        x_init = np.asarray(x_init, float)
        B, N = x_init.shape[0], self.N
        tr = np.asarray(traj_ref, float)
        x_ref = np.zeros((B, N + 1, 9)); x_ref[:, :, :4] = tr[:, :, :4]      # the pose reference rides in the first four columns
        batch = dict(x_init=x_init, x_ref=x_ref, u_ref=np.asarray(u_ref, float),
                     u_last=np.zeros((B, N, 5)) if u_last is None else np.asarray(u_last, float), u_guess=u_guess, x_guess=x_guess)
        c = self._cfg
        if c.n_obs:
            batch["circles"] = np.broadcast_to(self._circles_array(), (B, c.n_obs, 3)) if circles is None else circles
        out = self._backend_solve(batch, B)
        self.last_info = {k: out[k] for k in ("status", "iters", "kkt", "cost")}
        return out
