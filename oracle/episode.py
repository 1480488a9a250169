"""CPU restatement of Interface (interface_wholebody_qref.py, physical_sim=False) around the oracle solver.

TEST INFRASTRUCTURE ONLY.  One episode at a time, NumPy, line by line after the reference:
    run / timerCallback        :75-143     stateMachineUpdate   :146-228
    globalPlan2D               :247-266    globalPlanManipulator :277-297
    calcLocalRefTraj           :353-396    calcLocalRefPose      :398-410
MPCWholeBody.solve (:287-331) is oracle/solver.py (clip of x_init[6:] in place :290, U_last := previous U* :310),
setWeight (:119-139) swaps the diagonal weights, the `opti.subject_to(X[N,:2] == X_ref[N,:2])` of :167 is the solver's
flag bit 0 and stays set for the rest of the episode.  inverse_transformation: oracle/ik.py::solve_lm (parity
unpinned, see there)."""
import numpy as np

from mobile_manipulator_mpc_b200 import _abi
from . import ik as IK
from . import model as M
from . import solver as S

XLIM = np.array([[-100, -100, -np.inf, -2, -2, -np.pi, -np.pi / 2, -np.pi, 0],
                 [100, 100, np.inf, 2, 2, np.pi, np.pi / 2, 0, 3 * np.pi / 2]])
Q_DEFAULT = np.array([25.0, 25.0, 0.0, 0.0, 0.0, 5.0, 5.0, 5.0, 5.0])        # controllers/mpc_wholebody_qref.py:12-13
Q_ROTATE = np.array([5.0, 5.0, 5.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0])           # interface :176-178
Q_MANIPULATE = np.array([500.0, 500.0, 500.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0])  # interface :212-215


class Episode:
    def __init__(self, dt, t_move, t_manipulate, x_start, global_pose_target, circles, planes, N=20, mode=_abi.MODE_REFERENCE,
                 terminal_rows_on_sN=0):
        """terminal_rows_on_sN: 0 = the reference's NLP to the letter, 1 = the variant the GPU's
        reference mode solves (include/mmpc.h)."""
        self.dt, self.t_move, self.t_manip, self.N, self.mode = dt, t_move, t_manipulate, N, mode
        self.rows_on_sN = int(terminal_rows_on_sN)
        self.gp = np.asarray(global_pose_target, float)
        self.x_start = np.asarray(x_start, float).copy()
        wr = 0.6                                                                                     # :22
        self.x_target = np.array([self.gp[0] - wr * np.cos(self.gp[3]), self.gp[1] - wr * np.sin(self.gp[3]), self.gp[3],
                                  0, 0, 0, x_start[6], x_start[7], x_start[8]])                       # :23-31
        self.circles, self.planes = np.asarray(circles, float), np.asarray(planes, float)
        self.state = self.x_start.copy()
        self.flag = "move"
        self.traj_ref = None
        self.Qd = Q_DEFAULT.copy()
        self.term_eq = 0
        self.u_latest = np.zeros((N, 5))
        self.active = True
        self.steps = 0
        self.log = []

    # ---- plans and local references ----
    def _window(self, idx):
        idx = np.asarray(idx)
        min_d, min_i = 1e5, -1
        for i, xr in enumerate(self.traj_ref):                                                       # :364-372
            d = np.linalg.norm(self.state[idx] - xr[idx])
            if d < min_d:
                min_d, min_i = d, i
        rows = np.minimum(np.arange(min_i, min_i + self.N + 1), self.traj_ref.shape[0] - 1)          # :374-389
        self.local_ref = self.traj_ref[rows].copy()

    def _pose(self):
        self.local_ref = np.tile(self.traj_ref[-1], (self.N + 1, 1))                                 # :400
        self.local_ref[:, 2] = self.state[2] + M.angle_diff(self.traj_ref[-1, 2], self.state[2])     # :407-410

    def _update(self):
        x = self.state
        if self.flag == "move" and self.traj_ref is None:
            self.traj_ref = np.linspace(self.x_start, self.x_target, int(self.t_move / self.dt) + 1)  # :264
        if self.flag in ("move", "approach"):
            last = self.traj_ref[-1]
            if abs(x[0] - last[0]) <= 2 and abs(x[1] - last[1]) <= 2 and self.flag == "move":        # :153-167
                self.flag = "approach"
                self.term_eq = 1
            if np.linalg.norm(x[0:2] - last[0:2]) <= 0.2:                                            # :170-178
                self.flag = "rotate"
                self.Qd = Q_ROTATE.copy()
            elif self.flag == "move":
                self._window([0, 1])                                                                 # :188
            else:
                self._pose()
        if self.flag == "rotate":                                                                    # :192-197
            last = self.traj_ref[-1]
            if abs(M.angle_diff(x[2], last[2])) <= 0.5 * np.pi / 180 and np.linalg.norm(x[0:2] - last[0:2]) <= 0.01:
                self.flag = "move finish"
            else:
                self._pose()
        if self.flag == "move finish":                                                               # :204-216
            self.flag = "manipulate"
            lt = np.array([np.sqrt((self.gp[0] - x[0]) ** 2 + (self.gp[1] - x[1]) ** 2) + 0.007, 0.0, self.gp[2] - (0.606 + 0.333)])
            self.local_pose_target = lt
            q, st = IK.solve_lm(x[-3:], lt)
            if st != 0:
                self.flag = "ik failed"
                return False
            x_target = np.hstack((x[:6], q))                                                         # :284-287
            self.traj_ref = np.linspace(x, x_target, int(self.t_manip / self.dt) + 1)                # :293
            self.Qd = Q_MANIPULATE.copy()
        if self.flag == "manipulate":                                                                # :219-226
            e = np.asarray(M.forward_transformation(x)[0], float)[:3]
            if np.linalg.norm(e - self.gp[:3]) <= 0.01:
                self.flag = "manipulate finish"
                return False
            self._window([6, 7, 8])
        return True

    def step(self):
        """One timerCallback (:100-143).  Returns False once the episode is over."""
        if not self.active:
            return False
        self.steps += 1
        self.active = self._update()
        self.log.append((self.flag, self.state.copy()))
        if not self.active:
            return False
        self.state[6:] = np.clip(self.state[6:], XLIM[0, 6:], XLIM[1, 6:])                           # solve :290
        batch = dict(N=self.N, dt=self.dt, n_obs=self.circles.shape[0], n_pl=self.planes.shape[0], obs_per_stage=0,
                     x_init=self.state[None], x_ref=self.local_ref[None], u_ref=np.zeros((1, self.N, 5)),
                     u_last=self.u_latest[None], circles=self.circles[None], planes=self.planes[None],
                     n_pl_inst=np.array([self.planes.shape[0]], np.int32), flags=np.array([self.term_eq], np.uint8),
                     Qd=self.Qd, Pd=self.Qd)
        out = S.solve(batch, cfg=S.config_from_batch(batch, self.mode, terminal_rows_on_sN=self.rows_on_sN))
        self.status = int(out["status"][0])
        self.u_latest = out["U"][0].copy()                                                           # :330
        self.state = np.asarray(M.f_kinematics(self.state, self.u_latest[0], self.dt), float).reshape(9)  # :143
        return True

    def run(self, max_steps=1000):
        while self.active and self.steps < max_steps:
            self.step()
        return self.flag
