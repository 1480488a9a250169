"""One batched solve of BASELINE config 3 for ncu (host-sequenced rounds, so every kernel is a plain stream launch).
usage: ncu ... python scripts/ncu_driver.py [B] [reference|clean]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mode = _abi.MODE_CLEAN if (len(sys.argv) > 2 and sys.argv[2] == "clean") else _abi.MODE_REFERENCE
b = scenarios.make_batch(3, B)
S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B, mode=mode, kernel="staged_hostloop", max_iter=int(os.environ.get("MAXIT", "2000")))
o = S.solve_device(S.to_device(b)); torch.cuda.synchronize()
print("converged", float((o["status"] == 0).double().mean()), "launches", S.launch_count())
