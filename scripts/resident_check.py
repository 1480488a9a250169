"""The resident kernel (csrc/mmpc_resident.cu) against the staged solver on a B200: same results, latency of one instance,
throughput of small batches.  usage: python scripts/resident_check.py [out.json]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver


def timed(S, d, o, reps):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); S.solve_device(d, out=o); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


out = {}
for cfg, B, mode, mname in ((1, 1, _abi.MODE_REFERENCE, "reference"), (1, 1, _abi.MODE_CLEAN, "clean"),
                            (3, 148, _abi.MODE_REFERENCE, "reference"), (3, 1024, _abi.MODE_REFERENCE, "reference"),
                            (3, 1024, _abi.MODE_CLEAN, "clean"), (2, 512, _abi.MODE_REFERENCE, "reference")):
    batch = scenarios.make_batch(cfg, B)
    res = {}
    for kern in ("staged", "resident"):
        S = BatchSolver(N=batch["N"], dt=batch["dt"], n_obs=batch["n_obs"], n_pl=batch["n_pl"], B_max=B, mode=mode, kernel=kern)
        d = S.to_device(batch)
        o = S.solve_device(d); torch.cuda.synchronize()
        ms = timed(S, d, o, 20 if B <= 148 else 5)
        res[kern] = {k: v.cpu().numpy().copy() for k, v in o.items()}
        res[kern + "_ms"] = ms
        S.close()
    a, r = res["staged"], res["resident"]
    key = f"config{cfg}/B{B}/{mname}"
    both = (a["status"] == 0) & (r["status"] == 0)
    out[key] = dict(staged_ms=res["staged_ms"], resident_ms=res["resident_ms"],
                    status_equal=int((a["status"] == r["status"]).sum()), iters_equal=int((a["iters"] == r["iters"]).sum()),
                    U_bitwise=int((a["U"].reshape(B, -1) == r["U"].reshape(B, -1)).all(axis=1).sum()),
                    cost_max_rel=float(np.max(np.abs(a["cost"][both] - r["cost"][both]) / np.maximum(1, np.abs(a["cost"][both])))) if both.any() else None,
                    converged_staged=int((a["status"] == 0).sum()), converged_resident=int((r["status"] == 0).sum()),
                    mean_iters=float(r["iters"].mean()), B=B)
    print(key, out[key], flush=True)
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/resident_check.json"
os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
json.dump(out, open(path, "w"), indent=1)
