"""GPU check script: each kernel strategy on the BASELINE config shapes, timed on the device and
compared with the CPU oracle on a sample.  usage: first_gpu.py [lane,warp] [quick]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
from oracle import solver as osolver

print(torch.cuda.get_device_name(0))
KERNELS = sys.argv[1].split(",") if len(sys.argv) > 1 else ["lane", "warp"]
MODE = 0 if (len(sys.argv) > 3 and sys.argv[3] == "reference") else 1
CASES = ((1, 1, 1), (2, 4096, 256), (3, 8192, 256), (3, 65536, 64), (5, 2048, 32))
if len(sys.argv) > 2 and sys.argv[2] == "quick":
    CASES = ((1, 1, 1), (3, 8192, 128), (3, 65536, 32))
for kern in KERNELS:
    for cid, B, nchk in CASES:
        b = scenarios.make_batch(cid, B)
        S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B,
                        obs_per_stage=b["obs_per_stage"], kernel=kern, mode=MODE)
        print("kernel", kern, "config", cid, "B", B, S.occupancy(), flush=True)
        t = time.time(); o = S.solve_host(b); th = time.time() - t
        t = time.time(); o = S.solve_host(b); th2 = time.time() - t
        st = np.bincount(o["status"], minlength=6)
        print("  host path %.1f ms (2nd %.1f ms) status %s iters p50 %d p99 %d max %d mean %.1f" % (
            th * 1e3, th2 * 1e3, st, np.percentile(o["iters"], 50), np.percentile(o["iters"], 99), o["iters"].max(),
            o["iters"].mean()), flush=True)
        d = S.to_device(b)
        out = S.solve_device(d); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = S.solve_device(d, out=out); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        conv = int((out["status"] == 0).sum())
        print("  device path %.2f ms -> %.0f converged solves/s" % (ms, conv / ms * 1e3), flush=True)
        sub = {k: (v[:nchk] if isinstance(v, np.ndarray) else v) for k, v in b.items()}
        oc = osolver.solve(sub, mode=MODE, threads=8)
        both = (oc["status"] == 0) & (o["status"][:nchk] == 0)
        rel = np.abs(oc["cost"] - o["cost"][:nchk]) / np.abs(oc["cost"])
        du0 = np.abs(oc["U"][:, 0] - o["U"][:nchk, 0]).max(axis=1)
        print("  vs oracle: both converged %d/%d, cost rel max %.2e (>1e-5: %d), u0 abs max %.2e (>1e-4: %d), iters equal %.2f" % (
            both.sum(), nchk, rel[both].max(), (rel[both] > 1e-5).sum(), du0[both].max(), (du0[both] > 1e-4).sum(),
            (oc["iters"] == o["iters"][:nchk]).mean()), flush=True)
        S.close()
