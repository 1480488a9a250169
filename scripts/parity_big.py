"""GPU <-> oracle parity at a larger sample than the -m gpu tests use: exact counts for BASELINE configs 2, 3 and 5 (reference
NLP), staged solver and resident kernel.  usage: python scripts/parity_big.py [n_instances] [out.json]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
from oracle import solver as osolver, nlp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/parity_big.json"
rep = {}
for cid, B in ((3, n), (2, min(n, 4096)), (5, min(n, 1024))):
    b = scenarios.make_batch(cid, B)
    t = time.perf_counter(); ref = osolver.solve(b, mode=_abi.MODE_REFERENCE, threads=os.cpu_count() or 4); t_cpu = time.perf_counter() - t
    for kern in ("staged", "resident"):
        S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B, obs_per_stage=b["obs_per_stage"])
        try:
            S.set_kernel(kern)
        except Exception:
            S.close(); continue
        o = S.solve_host(b)
        both = (o["status"] == 0) & (ref["status"] == 0)
        rel = np.abs(o["cost"] - ref["cost"]) / np.maximum(np.abs(ref["cost"]), 1e-300)
        du0 = np.abs(o["U"][:, 0] - ref["U"][:, 0]).max(axis=1)
        outl = np.nonzero(both & ((rel > 1e-5) | (du0 > 1e-4)))[0]
        listed = []
        for i in outl[:40]:
            P = nlp.from_batch(b, int(i), "reference")
            va = P.violation(P.pack(o["X"][i], o["U"][i], o["s"][i])); vb = P.violation(P.pack(ref["X"][i], ref["U"][i], ref["s"][i]))
            listed.append(dict(instance=int(i), cost_gpu=float(o["cost"][i]), cost_oracle=float(ref["cost"][i]), du0=float(du0[i]),
                               violation_gpu=float(va), violation_oracle=float(vb), kkt_gpu=float(o["kkt"][i]), kkt_oracle=float(ref["kkt"][i])))
        good = both & ~np.isin(np.arange(B), outl)
        rep["config%d/%s" % (cid, kern)] = dict(
            B=B, nlp="reference", gpu_status_histogram=np.bincount(o["status"], minlength=6).tolist(),
            oracle_status_histogram=np.bincount(ref["status"], minlength=6).tolist(), status_differs=int((o["status"] != ref["status"]).sum()),
            both_converged=int(both.sum()), within_tolerance=int(good.sum()), outliers=int(len(outl)),
            outliers_both_feasible_kkt_points=int(sum(1 for l in listed if l["violation_gpu"] <= 1e-6 and l["violation_oracle"] <= 1e-6 and l["kkt_gpu"] <= 1e-8 and l["kkt_oracle"] <= 1e-8)),
            max_rel_cost_err_within=float(rel[good].max()) if good.any() else None, max_du0_within=float(du0[good].max()) if good.any() else None,
            iterations_equal=int((o["iters"] == ref["iters"])[both].sum()), max_violation_gpu_dyn=None, oracle_seconds=t_cpu, outlier_list=listed)
        print("config%d/%s" % (cid, kern), {k: v for k, v in rep["config%d/%s" % (cid, kern)].items() if k != "outlier_list"}, flush=True)
        S.close()
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(rep, open(out_path, "w"), indent=1)
