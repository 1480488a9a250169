"""Small solves for compute-sanitizer (memcheck / racecheck): B <= 64, both NLP variants, the CUDA-graph driver and the host loop,
the bulk ("fat") and the thin ("parts") kernels, the terminal-equality flag, moving circles, plus the helper kernels.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
for cid in (3, 5):
    b = scenarios.make_batch(cid, B if cid == 3 else 8)
    b["flags"] = (np.arange(b["x_init"].shape[0]) % 3 == 0).astype(np.uint8)
    for mode in (_abi.MODE_REFERENCE, _abi.MODE_CLEAN):
        for kern in ("staged", "staged_hostloop", "staged_fat"):
            S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=b["x_init"].shape[0], mode=mode, kernel=kern,
                            obs_per_stage=b["obs_per_stage"], max_iter=60)
            o = S.solve_host(b)
            print("config", cid, "mode", mode, kern, "status", np.bincount(o["status"], minlength=6).tolist(), flush=True)
            if kern == "staged" and cid == 3:
                d = S.to_device(b)
                x = d["x_init"]
                S.eval_model(x, None, d["circles"], d["planes"])
                U = torch.from_numpy(o["U"]).cuda()
                S.shift(U); S.plant_step(x, U[:, 0].contiguous())
                torch.cuda.synchronize()
            S.close()
print("sanitize_small done")
