"""A/B timing of library builds (MMPC_LIB=path): reference NLP, config 3, one context; graph-driven step time and per-phase device
times of one host-sequenced solve.  usage: MMPC_LIB=ab/lib_X.so python scripts/ab_time.py [B]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
b = scenarios.make_batch(3, B)
S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B)
d = S.to_device(b); o = S.solve_device(d); torch.cuda.synchronize()
ts = []
for _ in range(3):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); S.solve_device(d, out=o); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
S.set_profile(True); S.solve_device(d, out=o); torch.cuda.synchronize(); pm, pl, rounds = S.phase_times(); S.set_profile(False)
print(json.dumps(dict(lib=os.environ.get("MMPC_LIB", "default"), ms=min(ts), conv=float((o["status"] == 0).double().mean()),
                      phases={k: round(v, 1) for k, v in pm.items()})))
