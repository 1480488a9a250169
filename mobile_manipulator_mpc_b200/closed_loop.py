"""Device-side receding-horizon loop (BASELINE config 4): what Interface.timerCallback does
per MPC step without pybullet (interface_wholebody_qref.py:100-143) for B instances at once, with no host
round trip of the data:

    window      calcLocalRefTraj            :353-396   mmpc_window      (nearest row, rows [i*, i*+N])
    solve       controller.solve            :134       mmpc_solve       (u_last := previous U*, reference semantics
                                                                          controllers/mpc_wholebody_qref.py:295-310)
    warm start  (ours) shifted GUESS only                mmpc_shift       
    plant       current_state = f_dynamics  :143       mmpc_plant_step

Everything lives in torch CUDA tensors; the per-step host work is a handful of launches."""
import numpy as np
import torch

from . import _abi
from .batch_solver import BatchSolver


class ClosedLoop:
    def __init__(self, batch, x_glob, device=0, shift_guess=True, idx=(0, 1), **solver_kw):
        """``batch``: dict as scenarios.make_batch (x_init, circles, planes, n_pl_inst ...); ``x_glob``: global
        reference [M,9] or [B,M,9] (NumPy)."""
        self.N, self.B = batch["N"], int(batch["x_init"].shape[0])
        self.solver = BatchSolver(N=batch["N"], dt=batch["dt"], n_obs=batch["n_obs"], n_pl=batch["n_pl"], B_max=self.B,
                                  device=device, obs_per_stage=batch.get("obs_per_stage", 0), **solver_kw)
        dev = torch.device("cuda", device)
        self.dev = dev
        t = lambda a, dt=torch.float64: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
        self.x = t(batch["x_init"])
        self.x_glob = t(x_glob)
        self.static = dict(circles=t(batch.get("circles")), planes=t(batch.get("planes")),
                           n_pl_inst=t(batch.get("n_pl_inst"), torch.int32))
        self.u_last = torch.zeros((self.B, self.N, 5), dtype=torch.float64, device=dev)   # zeros on the first call (:298)
        self.u_guess = None
        self.shift_guess, self.idx = shift_guess, tuple(idx)
        self.out = None
        self.steps = 0
        self._conv = []
        self.failed = torch.zeros((), dtype=torch.int64, device=dev)
        c = self.solver.cfg
        self.q_lo = torch.tensor(list(c.xlim[0])[6:], dtype=torch.float64, device=dev)
        self.q_hi = torch.tensor(list(c.xlim[1])[6:], dtype=torch.float64, device=dev)

    def reset_counters(self):
        self._conv = []

    def conv_per_step(self):
        """converged instances of every counted step, as one device tensor (no host sync until it is read)"""
        return torch.stack(self._conv) if self._conv else torch.zeros(0, dtype=torch.int64, device=self.dev)

    def step(self, count=False):
        """One MPC step of every instance; returns (u0 [B,5], status [B]) as device tensors.  Nothing here waits for the
        device: the solve is one CUDA-graph launch, so a host thread can keep several sub-batches in flight."""
        S = self.solver
        x_ref, u_ref = S.window(self.x, self.x_glob, None, self.idx)
        inp = dict(x_init=self.x, x_ref=x_ref, u_ref=u_ref, u_last=self.u_last, u_guess=self.u_guess)
        inp.update({k: v for k, v in self.static.items() if v is not None})
        self.out = S.solve_device(inp, out=self.out)
        U = self.out["U"]
        st = self.out["status"]
        ok = (st != _abi.STATUS_NAN) & (st != _abi.STATUS_FACTOR)   # an iterate exists (converged, or the best point of a stalled search)
        u0 = U[:, 0, :].contiguous()
        # solve() clips the caller's x_init[6:] IN PLACE (controllers/mpc_wholebody_qref.py:290): the plant sees the clipped joints
        x_in = self.x.clone()
        x_in[:, 6:] = torch.minimum(torch.maximum(x_in[:, 6:], self.q_lo), self.q_hi)
        xn = S.plant_step(x_in, u0)
        # A failed solve is fatal in the reference (:329).  A batch cannot die: a solve without a finite iterate (NaN,
        # factorisation failure) holds the instance's state and U_last for this step; a stalled one (line search, iteration
        # cap) applies its last iterate.  ``failed`` counts every solve that is not converged / acceptable.
        self.x = torch.where(ok[:, None], xn, x_in)
        self.u_last = torch.where(ok[:, None, None], U, self.u_last)   # U_last := previous U*, same index (:310)
        self.failed = self.failed + ((st != _abi.STATUS_CONVERGED) & (st != _abi.STATUS_ACCEPTABLE)).sum()
        self.u_guess = S.shift(self.u_last) if self.shift_guess else None   # only the GUESS is shifted
        self.steps += 1
        if count:
            self._conv.append((self.out["status"] == _abi.STATUS_CONVERGED).sum())
        return u0, self.out["status"]

    def run(self, steps):
        conv = 0
        for _ in range(steps):
            _, st = self.step()
            conv += int((st == _abi.STATUS_CONVERGED).sum())
        return conv


def config4(B, seed=4, t_move=50.0, dt=0.1, N=20):
    """BASELINE config 4 workload: config 3's instances (16 random circles, default_rng(4)),
    every instance's global reference = linspace from its start to the base goal stretched to t_move/dt + 1
    rows so that the window keeps moving for 500 steps."""
    from . import scenarios
    b = scenarios.make_batch(3, B, N=N, seed=seed)
    x0 = b["x_init"]
    goal = np.array([5.0, 5.0, -np.pi, 0, 0, 0])
    M = int(round(t_move / dt)) + 1
    tgt = np.concatenate([np.tile(goal, (B, 1)), x0[:, 6:]], axis=1)
    w = np.linspace(0.0, 1.0, M)[None, :, None]
    x_glob = x0[:, None, :] * (1 - w) + tgt[:, None, :] * w
    return b, np.ascontiguousarray(x_glob)
