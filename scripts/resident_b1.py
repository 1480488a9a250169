"""One instance (config 1) and one wave (config 3, B = SM count) through the resident kernel: latency.  With an MMPC_RES_PROF
build (MMPC_LIB=ab/lib_PROF.so) the kernel prints block 0's cycles per phase.  usage: python scripts/resident_b1.py [kernel]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
kern = sys.argv[1] if len(sys.argv) > 1 else "resident"
for cfg, B in ((1, 1), (3, 148), (3, 296), (3, 592), (3, 1024), (3, 2048), (2, 1024)):
    b = scenarios.make_batch(cfg, B)
    S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B, kernel=kern)
    d = S.to_device(b); o = S.solve_device(d); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); S.solve_device(d, out=o); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    print(json.dumps(dict(lib=os.environ.get("MMPC_LIB", "default"), kernel=kern, config=cfg, B=B, p50_ms=float(np.median(ts)), min_ms=min(ts),
                          iters=float(o["iters"].double().mean()), conv=float((o["status"] == 0).double().mean()))), flush=True)
    S.close()
