import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mobile_manipulator_mpc_b200 import scenarios
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
cid = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kern = sys.argv[4] if len(sys.argv) > 4 else "auto"
b = scenarios.make_batch(cid, B)
from mobile_manipulator_mpc_b200 import _abi
mode = _abi.MODE_REFERENCE if os.environ.get("MMPC_MODE") == "reference" else _abi.MODE_CLEAN   # reference: the literal NLP (quirks 1-3)
S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B, obs_per_stage=b["obs_per_stage"], kernel=kern, mode=mode)
d = S.to_device(b)
out = None
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = S.solve_device(d, out=out); e1.record(); torch.cuda.synchronize()
    print("rep", r, "ms", e0.elapsed_time(e1), "converged", int((out["status"] == 0).sum()), "iters sum", int(out["iters"].sum()))
if os.environ.get("MMPC_PHASES"):
    S.set_profile(True)
    out = S.solve_device(d, out=out); torch.cuda.synchronize()
    pm, pl, rounds = S.phase_times()
    tot = sum(pm.values())
    print("phases (ms, launches) over %d rounds, total %.1f ms:" % (rounds, tot),
          " ".join("%s=%.1f/%d" % (k, pm[k], pl[k]) for k in pm if pl[k]))
