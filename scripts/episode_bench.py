"""Throughput of whole task episodes on one B200: B robots run move -> approach -> rotate ->
manipulate -> finish in lock step on the device (mobile_manipulator_mpc_b200/episodes.py).
usage: python scripts/episode_bench.py [B] [reference|clean] [max_iter] [max_steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mobile_manipulator_mpc_b200 import scenarios, _abi
from mobile_manipulator_mpc_b200.episodes import BatchedInterface

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = _abi.MODE_CLEAN if (len(sys.argv) > 2 and sys.argv[2] == "clean") else _abi.MODE_REFERENCE
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 300
max_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 400
xs, gps, circ, pls, npl = scenarios.episode_batch(B)
T = BatchedInterface(0.1, 5, 2, xs, gps, circ, pls, npl, N=20, mode=mode, max_iter=max_iter)
T.step(); torch.cuda.synchronize()            # warm-up step (allocates the solver workspaces), then restart
T.close()
T = BatchedInterface(0.1, 5, 2, xs, gps, circ, pls, npl, N=20, mode=mode, max_iter=max_iter)
lat = []; hist = []
t0 = time.time()
e0 = torch.cuda.Event(enable_timing=True); e0.record()
while T.steps < max_steps:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); n = T.step(); b.record(); b.synchronize()
    lat.append(a.elapsed_time(b)); hist.append(n)
    if n == 0:
        break
e1 = torch.cuda.Event(enable_timing=True); e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1)
task = T.task.cpu().numpy()
print(json.dumps(dict(metric="episode instance-steps/s", B=B, mode="reference" if mode == _abi.MODE_REFERENCE else "clean",
                      max_iter=max_iter, steps=T.steps, instance_steps=T.solves, device_ms=ms, wall_s=time.time() - t0,
                      instance_steps_per_s=T.solves / (ms * 1e-3), episodes_per_s=B / (ms * 1e-3),
                      p50_step_ms=float(np.median(lat)), p99_step_ms=float(np.percentile(lat, 99)),
                      nonconverged_fraction=T.nonconverged / max(T.solves, 1),
                      finished=int((task == _abi.TASK_FINISHED).sum()), ik_failed=int((task == _abi.TASK_IK_FAILED).sum()),
                      still_running=int((task < _abi.TASK_FINISHED).sum()),
                      active_per_step_p50=float(np.median(hist)))))
