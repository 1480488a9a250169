"""Throughput of whole task episodes on one B200: B robots run move -> approach -> rotate ->
manipulate -> finish in lock step on the device (mobile_manipulator_mpc_b200/episodes.py).
usage: python scripts/episode_bench.py [B] [reference|clean] [max_iter] [max_steps] [shards]"""
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
shards = int(sys.argv[5]) if len(sys.argv) > 5 else 4
from mobile_manipulator_mpc_b200.sharding import run_interleaved


def make():
    """K independent sub-batches, each with its own solver contexts, host thread and stream (sharding.run_interleaved)"""
    out = []
    for i in range(shards):
        sl = slice(i * B // shards, (i + 1) * B // shards)
        out.append(BatchedInterface(0.1, 5, 2, xs[sl], gps[sl], circ[sl], pls[sl], npl[sl], N=20, mode=mode, max_iter=max_iter))
    return out


def step(T):
    if getattr(T, "done", False):
        return 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); n = T.step(); b.record(); b.synchronize()
    T.lat.append(a.elapsed_time(b)); T.done = n == 0
    return n


Ts = make()
for T in Ts:
    T.lat = []
run_interleaved(Ts, step, 1); torch.cuda.synchronize()   # warm-up step (allocates the solver workspaces), then restart
for T in Ts:
    T.close()
Ts = make()
for T in Ts:
    T.lat = []
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
hist = run_interleaved(Ts, step, max_steps)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1)
task = np.concatenate([T.task.cpu().numpy() for T in Ts])
lat = [v for T in Ts for v in T.lat]
solves = sum(T.solves for T in Ts); nonconv = sum(T.nonconverged for T in Ts)
print(json.dumps(dict(metric="episode instance-steps/s", B=B, shards=shards, mode="reference" if mode == _abi.MODE_REFERENCE else "clean",
                      max_iter=max_iter, steps=max(T.steps for T in Ts), instance_steps=solves, device_ms=ms, wall_s=time.time() - t0,
                      instance_steps_per_s=solves / (ms * 1e-3), episodes_per_s=B / (ms * 1e-3),
                      p50_shard_step_ms=float(np.median(lat)), p99_shard_step_ms=float(np.percentile(lat, 99)),
                      nonconverged_fraction=nonconv / max(solves, 1),
                      finished=int((task == _abi.TASK_FINISHED).sum()), ik_failed=int((task == _abi.TASK_IK_FAILED).sum()),
                      still_running=int((task < _abi.TASK_FINISHED).sum()))))
