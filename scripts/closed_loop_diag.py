"""Iteration statistics of the closed loop (BASELINE config 4 workload), one sub-batch: per step the mean / p99 / max iteration count
(the max is the number of rounds the staged solve runs) and the solve time.  usage: python scripts/closed_loop_diag.py [B] [steps]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobile_manipulator_mpc_b200 import closed_loop
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5461
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
b, xg = closed_loop.config4(B)
L = closed_loop.ClosedLoop(b, xg, shift_guess=True)
rows = []
for s in range(steps):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); L.step(); e.record(); torch.cuda.synchronize()
    it = L.out["iters"].cpu().numpy(); st = L.out["status"].cpu().numpy()
    rows.append(dict(step=s, ms=round(a.elapsed_time(e), 1), mean=round(float(it.mean()), 1), p99=int(np.percentile(it, 99)), max=int(it.max()),
                     nonconv=int((st != 0).sum())))
for r in rows[:3] + rows[10:13] + rows[-5:]:
    print(r)
ms = np.array([r["ms"] for r in rows[3:]]); mx = np.array([r["max"] for r in rows[3:]]); mean = np.array([r["mean"] for r in rows[3:]])
print(json.dumps(dict(B=B, steps=steps, ms_mean=float(ms.mean()), iters_mean=float(mean.mean()), max_iters_mean=float(mx.mean()), max_iters_p90=float(np.percentile(mx, 90)))))
