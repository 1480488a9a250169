import os, sys, json
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
B = 65536
b = scenarios.make_batch(3, B)
S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B, mode=_abi.MODE_CLEAN)
d = S.to_device(b); o = S.solve_device(d); torch.cuda.synchronize()
ts = []
for _ in range(3):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); S.solve_device(d, out=o); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
S.set_profile(True); S.solve_device(d, out=o); torch.cuda.synchronize(); pm, pl, rounds = S.phase_times(); S.set_profile(False)
print(json.dumps(dict(tiles=os.environ.get("MMPC_PARTS_TILES"), ms=min(ts), conv=float((o["status"] == 0).double().mean()), phases={k: round(v, 1) for k, v in pm.items()})))
