"""First look at the CUDA-graph driver on a B200: time one batched solve (reference NLP and clean NLP), graph against host loop,
single context; B = 1 latency.  usage: python scripts/graph_check.py [B]"""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
batch = scenarios.make_batch(3, B)
out = {}
for mode, mname in ((_abi.MODE_REFERENCE, "reference"), (_abi.MODE_CLEAN, "clean")):
    for kern in ("staged", "staged_hostloop"):
        S = BatchSolver(N=batch["N"], dt=batch["dt"], n_obs=batch["n_obs"], n_pl=batch["n_pl"], B_max=B, mode=mode, kernel=kern)
        d = S.to_device(batch)
        o = S.solve_device(d); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); S.solve_device(d, out=o); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        conv = int((o["status"] == 0).sum())
        st = np.bincount(o["status"].cpu().numpy(), minlength=6).tolist()
        out[f"{mname}/{kern}"] = dict(ms=min(ts), solves_per_s=conv / (min(ts) * 1e-3), converged=conv / B, status_hist=st,
                                      mean_iters=float(o["iters"].double().mean()), launches=S.launch_count())
        print(mname, kern, out[f"{mname}/{kern}"], flush=True)
        S.close()
b1 = scenarios.make_batch(1, 1)
for kern in ("staged", "staged_hostloop"):
    S1 = BatchSolver(N=b1["N"], dt=b1["dt"], n_obs=b1["n_obs"], n_pl=b1["n_pl"], B_max=1, mode=_abi.MODE_REFERENCE, kernel=kern)
    d1 = S1.to_device(b1); o1 = S1.solve_device(d1); torch.cuda.synchronize()
    l1 = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); S1.solve_device(d1, out=o1); b.record(); torch.cuda.synchronize()
        l1.append(a.elapsed_time(b))
    out[f"B1/{kern}"] = dict(p50_ms=float(np.median(l1)), iters=int(o1["iters"][0]))
    print("B=1", kern, out[f"B1/{kern}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/graph_check.json", "w"), indent=1)
