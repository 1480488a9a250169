"""Batched episodes on the GPU against the CPU restatement (oracle/episode.py), step by step.
usage: python scripts/episode_check.py [B] [reference|clean] [n_checked] [free|forced] [on_sN]
The oracle and the GPU solve the same variant of the terminal self-collision rows: the reference to the
letter (rows on s[N-1]) unless "on_sN" is given."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.episodes import BatchedInterface
from oracle.episode import Episode

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = _abi.MODE_CLEAN if (len(sys.argv) > 2 and sys.argv[2] == "clean") else _abi.MODE_REFERENCE
ncheck = int(sys.argv[3]) if len(sys.argv) > 3 else min(B, 4)
forced = len(sys.argv) > 4 and sys.argv[4] == "forced"   # re-seed the oracle with the GPU's state and U* after every step
literal = not (len(sys.argv) > 5 and sys.argv[5] == "on_sN")  # terminal self-collision rows on s[N-1] (the reference to the letter) or on s[N]
xs, gps, circ, pls, npl = scenarios.episode_batch(B)
T = BatchedInterface(0.1, 5, 2, xs, gps, circ, pls, npl, N=20, mode=mode, terminal_rows_on_sN=0 if literal else 1)
eps = [Episode(0.1, 5, 2, xs[b], gps[b], circ[b], pls[b][:npl[b]], N=20, mode=mode, terminal_rows_on_sN=0 if literal else 1) for b in range(ncheck)]
t0 = time.time(); worst = 0.0; mism = 0
while T.steps < 600:
    x_prev = T.x.cpu().numpy().copy(); ul_prev = T.u_last.cpu().numpy().copy()
    n = T.step()
    task = T.task.cpu().numpy(); x = T.x.cpu().numpy(); stat = T.status.cpu().numpy()
    for b, ep in enumerate(eps):
        if ep.active:
            ep.step()
        name = _abi.TASK_NAMES[task[b]]
        if name != ep.flag:
            mism += 1
            print("step", T.steps, "episode", b, "task", name, "oracle", ep.flag)
        d = float(np.abs(x[b] - ep.state).max())
        if (stat[b] != 0 and task[b] < 5) or (d > 1e-6 and worst <= 1e-6) or d > 1e-5:
            print("step", T.steps, "episode", b, name, "gpu status", stat[b], "oracle status", getattr(ep, "status", None), "|dx| %.3e" % d)
        worst = max(worst, d)
        if d > 1e-3 and os.environ.get("EPISODE_DUMP"):
            np.savez(os.path.join(os.environ["EPISODE_DUMP"], "episode_dump_b%d_s%d.npz" % (b, T.steps)), x_init=x_prev[b],
                     x_ref=T.x_ref[b].cpu().numpy(), u_last=ul_prev[b], flags=T.flags[b].cpu().numpy(), wset=T.wset[b].cpu().numpy(),
                     circles=circ[b], planes=pls[b], npl=npl[b], U_gpu=T.u_last[b].cpu().numpy(), U_oracle=ep.u_latest, mode=mode)
        if forced and task[b] < 5:
            ep.state = x[b].copy(); ep.u_latest = T.u_last[b].cpu().numpy().copy()
    if n == 0:
        break
torch.cuda.synchronize()
print("B", B, "steps", T.steps, "solves", T.solves, "nonconverged", T.nonconverged, "wall %.1f s" % (time.time() - t0))
print("final tasks", np.bincount(T.task.cpu().numpy(), minlength=7), "flag mismatches", mism, "max |x - oracle|", worst)
