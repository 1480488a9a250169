"""Throughput with T solver contexts running concurrently (one host thread + stream + workspace each):
the thin tail of one batch overlaps the bulk of the next.  usage: overlap_test.py [B] [steps] [T...]"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mobile_manipulator_mpc_b200 import scenarios
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
Ts = [int(a) for a in sys.argv[3:]] or [1, 2, 3]
b = scenarios.make_batch(3, B)
for T in Ts:
    ctx = []
    for t in range(T):
        S = BatchSolver(N=b["N"], dt=b["dt"], n_obs=b["n_obs"], n_pl=b["n_pl"], B_max=B)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            d = S.to_device(b); out = S.solve_device(d)
        ctx.append((S, st, d, out))
    torch.cuda.synchronize()
    def work(t):
        S, st, d, out = ctx[t]
        with torch.cuda.stream(st):
            for i in range(t, steps, T):
                S.solve_device(d, out=out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    [x.start() for x in th]; [x.join() for x in th]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    conv = int((ctx[0][3]["status"] == 0).sum())
    print("contexts %d: %d steps of B=%d in %.1f ms -> %.0f converged solves/s" % (T, steps, B, dt * 1e3, conv * steps / dt), flush=True)
    for S, *_ in ctx: S.close()
    del ctx; torch.cuda.empty_cache()
