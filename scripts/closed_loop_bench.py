"""BASELINE config 4: closed-loop receding-horizon rollout, everything on device.
usage: closed_loop_bench.py [B] [steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mobile_manipulator_mpc_b200 import closed_loop
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
b, x_glob = closed_loop.config4(B)
for shift in (True, False):
    L = closed_loop.ClosedLoop(b, x_glob, shift_guess=shift)
    L.run(2); torch.cuda.synchronize()
    t0 = time.perf_counter(); lat = []; conv = 0; iters = 0
    for i in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _, st = L.step(); e1.record(); torch.cuda.synchronize()
        lat.append(e0.elapsed_time(e1)); conv += int((st == 0).sum()); iters += int(L.out["iters"].sum())
    dt = time.perf_counter() - t0
    d = torch.linalg.norm(L.x[:, :2] - torch.tensor([5.0, 5.0], device=L.x.device, dtype=torch.float64), dim=1)
    print("shift_guess=%s: B=%d, %d closed-loop steps in %.2f s -> %.0f instance-steps/s, p50 step latency %.1f ms, converged %.4f, "
          "mean iterations %.1f, mean distance to goal %.2f m" % (shift, B, steps, dt, conv / dt, np.median(lat), conv / (B * steps),
                                                                 iters / (B * steps), float(d.mean())), flush=True)
    L.solver.close()

# ---- K interleaved shards of the same total batch: the tail of one shard's solve overlaps the bulk of another's ----
from mobile_manipulator_mpc_b200.sharding import run_interleaved
for K in (2, 4):
    subs = []
    for i in range(K):
        sl = slice(i * B // K, (i + 1) * B // K)
        bi = {k: (v[sl] if isinstance(v, np.ndarray) else v) for k, v in b.items()}
        subs.append(closed_loop.ClosedLoop(bi, x_glob[sl], shift_guess=True))
    run_interleaved(subs, lambda L: L.step(), 2); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = run_interleaved(subs, lambda L: int((L.step()[1] == 0).sum()), steps)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); conv = sum(sum(r) for r in res)
    print("%d interleaved shards of %d: %d closed-loop steps in %.2f s (device) -> %.0f instance-steps/s, %.1f ms per step of the whole batch, "
          "converged %.4f" % (K, B // K, steps, ms * 1e-3, conv / (ms * 1e-3), ms / steps, conv / (B * steps)), flush=True)
    for L in subs:
        L.solver.close()
