"""One resident-kernel solve of BASELINE config 3 for ncu.  usage: ncu -k regex:resident ... python scripts/ncu_resident.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobile_manipulator_mpc_b200 import scenarios
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
b = scenarios.make_batch(3, B)
S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B, kernel="resident")
o = S.solve_device(S.to_device(b)); torch.cuda.synchronize()
print("converged", float((o["status"] == 0).double().mean()), "iters", float(o["iters"].double().sum()), "launches", S.launch_count())
