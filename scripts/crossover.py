import os, sys, json
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
from mobile_manipulator_mpc_b200 import scenarios
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
for cid, B in ((3, 148), (3, 296), (3, 592), (3, 1024), (3, 2048), (2, 592), (2, 1480), (2, 2960), (5, 148), (5, 444), (5, 1480)):
    b = scenarios.make_batch(cid, B)
    r = {}
    for kern in ("resident", "staged"):
        S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B, kernel=kern, obs_per_stage=b["obs_per_stage"])
        d = S.to_device(b); o = S.solve_device(d); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); S.solve_device(d, out=o); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
        r[kern] = round(float(np.median(ts)), 2); S.close()
    print(cid, B, r, flush=True)
