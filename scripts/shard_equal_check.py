"""2-GPU check of SURVEY.md section 7 (iii): the sharded result of a 1,024-instance batch equals the single-GPU result bit for bit.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/shard_equal_check.py
Every rank solves its contiguous shard on its own GPU, NCCL all-gathers U / cost / status / iters, rank 0 solves the whole batch on
its GPU and compares."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from mobile_manipulator_mpc_b200 import scenarios, sharding, _abi
from mobile_manipulator_mpc_b200.batch_solver import BatchSolver

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 1024
res = {}
# (the clean NLP's thin rounds use the warp-specialised part kernels, which sum in another order than the bulk kernels and are
# entered at a round that depends on the batch size: bitwise equality there needs kernel="staged_fat")
for mode, name, kern in ((_abi.MODE_REFERENCE, "reference", "staged"), (_abi.MODE_CLEAN, "clean_fat", "staged_fat")):
    batch = scenarios.make_batch(3, B)
    lo, hi = sharding.shard_bounds(B, world, rank)
    sub = {k: (v[lo:hi] if isinstance(v, np.ndarray) else v) for k, v in batch.items()}
    S = BatchSolver(N=batch["N"], dt=batch["dt"], n_obs=batch["n_obs"], n_pl=batch["n_pl"], B_max=B, device=local, mode=mode, kernel=kern)
    o = S.solve_device(S.to_device(sub))
    gathered = {}
    for k in ("U", "cost", "status", "iters"):
        t = o[k].contiguous()
        full = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t)
        gathered[k] = full
    if rank == 0:
        w = S.solve_device(S.to_device(batch))
        torch.cuda.synchronize()
        eq = {k: bool(torch.equal(gathered[k], w[k])) for k in gathered}
        res[name] = dict(equal=eq, converged=int((w["status"] == 0).sum()), B=B, world=world)
        assert all(eq.values()), (name, eq)
    S.close()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/r2_shard_equal_%dgpu.json" % world, "w"), indent=1)
    print("sharded == single-GPU, bit for bit:", res)
dist.barrier(); dist.destroy_process_group()
