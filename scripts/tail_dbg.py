import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
B = int(sys.argv[1]); kern = sys.argv[2]
b = scenarios.make_batch(3, B)
S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B, kernel=kern)
d = S.to_device(b)
print("start", B, kern, os.environ.get("MMPC_RESIDENT_TAIL"), flush=True)
t = time.time(); o = S.solve_device(d); print("launched %.3f s" % (time.time() - t), flush=True); torch.cuda.synchronize(); print("solve 1 done %.3f s" % (time.time() - t), "conv", float((o["status"] == 0).double().mean()), "launches", S.launch_count(), flush=True)
t = time.time(); o = S.solve_device(d, out=o); torch.cuda.synchronize(); print("solve 2 done %.3f s" % (time.time() - t), flush=True)
S.close(); print("closed", flush=True)
