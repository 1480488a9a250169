"""Batched task episodes: interface_wholebody_qref.Interface (physical_sim=False) for B robots at once, on the device
(SURVEY.md 8(f) rows 2 and 3).

Per MPC step (Interface.timerCallback :100-143):

    stateMachineUpdate :146-228   mmpc_episode_update   one kernel: transitions move -> approach -> rotate -> manipulate ->
                                                        finish, the local reference of the step (calcLocalRefTraj :353-396 /
                                                        calcLocalRefPose :398-410), the terminal-equality flag (:167), the
                                                        weight set, and at the hand-off the IK (manipulator_3DoF.py:79-133) and
                                                        the joint-space plan (:277-297)
    controller.solve   :134       mmpc_solve            the three weight sets of the machine (constructor defaults, 'rotate'
                                                        :176-178, 'manipulate' :212-215) are three solver contexts; every step
                                                        the active episodes are gathered by weight set, solved, scattered back
    f_dynamics         :143       mmpc_plant_step

States, references, flags and U_last stay in HBM; the host sees three group sizes per step."""
import numpy as np
import torch

from . import _abi
from .batch_solver import BatchSolver

# P = Q diagonals of the machine's three setWeight states
WEIGHT_SETS = (np.array([25.0, 25.0, 0.0, 0.0, 0.0, 5.0, 5.0, 5.0, 5.0]),        # controllers/mpc_wholebody_qref.py:12-13
               np.array([5.0, 5.0, 5.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0]),          # interface_wholebody_qref.py:176-178
               np.array([500.0, 500.0, 500.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0]))    # interface_wholebody_qref.py:212-215
WORKING_RADIUS = 0.6  # interface_wholebody_qref.py:22


class BatchedInterface:
    def __init__(self, dt, t_move, t_manipulate, x_start, global_pose_target, circles, planes, n_pl_inst=None, N=20,
                 mode=_abi.MODE_REFERENCE, device=0, max_iter=2000, terminal_rows_on_sN=0):
        """x_start [B,9], global_pose_target [B,4] (x y z psi), circles [B,n_obs,3], planes [B,n_pl,6] (NumPy).
        max_iter: 'ipopt.max_iter' (2000 in the reference, controllers/mpc_wholebody_qref.py:280).  One instance that runs
        into the cap holds up its whole group for that many rounds, so throughput runs lower it."""
        x_start = np.ascontiguousarray(x_start, dtype=np.float64)
        gp = np.ascontiguousarray(global_pose_target, dtype=np.float64)
        self.B, self.N, self.dt = int(x_start.shape[0]), int(N), float(dt)
        B = self.B
        self.n_move, self.n_manip = int(t_move / dt), int(t_manipulate / dt)
        self.M = max(self.n_move, self.n_manip) + 1
        dev = torch.device("cuda", device)
        self.dev = dev
        kw = dict(N=N, dt=dt, n_obs=int(circles.shape[1]), n_pl=int(planes.shape[1]), B_max=B, device=device, mode=mode,
                  max_iter=int(max_iter), terminal_rows_on_sN=int(terminal_rows_on_sN))
        self.solvers = []
        for w in WEIGHT_SETS:
            s = BatchSolver(**kw)
            s.set_weights(Q=np.diag(w), P=np.diag(w))
            self.solvers.append(s)
        # globalPlan2D :247-266 (NumPy's own linspace, so the rows are the reference's bit for bit)
        x_target = np.stack([gp[:, 0] - WORKING_RADIUS * np.cos(gp[:, 3]), gp[:, 1] - WORKING_RADIUS * np.sin(gp[:, 3]), gp[:, 3],
                             np.zeros(B), np.zeros(B), np.zeros(B), x_start[:, 6], x_start[:, 7], x_start[:, 8]], axis=1)  # :23-31
        traj = np.zeros((B, self.M, 9))
        for b in range(B):
            traj[b, :self.n_move + 1] = np.linspace(x_start[b], x_target[b], self.n_move + 1)
        t = lambda a, dt_=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt_)
        self.x = t(x_start)
        self.pose_target = t(gp)
        self.traj = t(traj)
        self.traj_len = torch.full((B,), self.n_move + 1, dtype=torch.int32, device=dev)
        self.task = torch.zeros(B, dtype=torch.int32, device=dev)
        self.flags = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.wset = torch.zeros(B, dtype=torch.int32, device=dev)
        self.active = torch.ones(B, dtype=torch.int32, device=dev)
        self.x_ref = torch.zeros((B, N + 1, 9), dtype=torch.float64, device=dev)
        self.u_ref = torch.zeros((B, N, 5), dtype=torch.float64, device=dev)
        self.local_pose_target = torch.zeros((B, 3), dtype=torch.float64, device=dev)
        self.ik_status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.u_last = torch.zeros((B, N, 5), dtype=torch.float64, device=dev)   # zeros before the first solve (:298)
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.circles, self.planes = t(circles), t(planes)
        self.n_pl_inst = t(n_pl_inst if n_pl_inst is not None else np.full(B, planes.shape[1]), torch.int32)
        xlim = np.array([[-np.pi / 2, -np.pi, 0.0], [np.pi / 2, 0.0, 3 * np.pi / 2]])  # controllers/mpc_wholebody_qref.py:19-21
        self.q_lo, self.q_hi = t(xlim[0]), t(xlim[1])
        self.steps = 0
        self.solves = 0
        self.nonconverged = 0

    def close(self):
        for s in self.solvers:
            s.close()

    def update(self):
        """stateMachineUpdate of every episode; afterwards task / active / wset / x_ref / flags describe this step."""
        io = dict(x=self.x, pose_target=self.pose_target, traj=self.traj, traj_len=self.traj_len, task=self.task,
                  flags=self.flags, wset=self.wset, active=self.active, x_ref=self.x_ref, u_ref=self.u_ref,
                  local_pose_target=self.local_pose_target, ik_status=self.ik_status)
        self.solvers[0].episode_update(io, self.B, self.M, self.n_manip)

    def step(self):
        """One timerCallback of every episode that is still running; returns the number of episodes solved."""
        self.update()
        act = self.active.bool()
        # solve() clips the caller's x_init[6:] in place (controllers/mpc_wholebody_qref.py:290); the plant then sees it
        self.x[:, 6:] = torch.where(act[:, None], torch.minimum(torch.maximum(self.x[:, 6:], self.q_lo), self.q_hi), self.x[:, 6:])
        n_solved = 0
        for w, S in enumerate(self.solvers):
            idx = torch.nonzero(act & (self.wset == w)).flatten()
            n = int(idx.numel())   # the one host sync per group
            if n == 0:
                continue
            g = lambda a: a.index_select(0, idx).contiguous()
            inp = dict(x_init=g(self.x), x_ref=g(self.x_ref), u_ref=g(self.u_ref), u_last=g(self.u_last), circles=g(self.circles),
                       planes=g(self.planes), n_pl_inst=g(self.n_pl_inst), flags=g(self.flags))
            out = S.solve_device(inp, want=())
            U = out["U"]
            st = out["status"]
            ok = (st != _abi.STATUS_NAN) & (st != _abi.STATUS_FACTOR)   # an iterate exists (converged, or the best point of a stalled search)
            # A failed solve is fatal in the reference (controllers/mpc_wholebody_qref.py:329).  A batch cannot die: a solve that
            # ends without a finite iterate (NaN, factorisation failure) leaves the episode where it is for this step; one that
            # stalled (line search, iteration cap) applies its last iterate, like the restated Interface (oracle/episode.py).
            # Both are counted in ``nonconverged``.
            self.u_last.index_copy_(0, idx, torch.where(ok[:, None, None], U, inp["u_last"]))   # U_last := previous U* (:310, :330)
            self.status.index_copy_(0, idx, st)
            xn = S.plant_step(inp["x_init"], U[:, 0, :].contiguous())                  # :143
            self.x.index_copy_(0, idx, torch.where(ok[:, None], xn, inp["x_init"]))
            n_solved += n
            self.nonconverged += int(((st != _abi.STATUS_CONVERGED) & (st != _abi.STATUS_ACCEPTABLE)).sum())
        self.steps += 1
        self.solves += n_solved
        return n_solved

    def run(self, max_steps=1000, log=None):
        """Until every episode is over or max_steps; ``log``: optional list that receives (task, x) NumPy copies per step."""
        while self.steps < max_steps:
            n = self.step()
            if log is not None:
                log.append((self.task.cpu().numpy().copy(), self.x.cpu().numpy().copy()))
            if n == 0:
                break
        return self.task.cpu().numpy()
