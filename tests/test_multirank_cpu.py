"""CPU tier: the N>1 host logic (weak-scaling shards + result gather) with world_size 2 on gloo."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mobile_manipulator_mpc_b200 import sharding, scenarios
    batch = scenarios.make_batch(3, 32, seed=sharding.rank_seed(3, rank))
    # stand-in for the device solve: any deterministic per-instance function of the inputs
    u0 = torch.from_numpy(batch["x_init"][:, :5] * 2.0)
    status = torch.from_numpy((batch["n_pl_inst"] == 3).astype(np.int32))
    allu0, allst = sharding.gather_results(dist, u0, status)
    stats = sharding.reduce_stats(dist, converged=int((status == 0).sum()), seconds=0.1 * (rank + 1))
    q.put((rank, allu0.numpy(), allst.numpy(), stats, batch["x_init"]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_order_and_stats():
    world, port = 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    x0 = np.concatenate([r[4] for r in res])
    assert not np.array_equal(res[0][4], res[1][4])             # ranks own different shards
    for r in res:
        assert r[1].shape == (64, 5) and np.array_equal(r[1], x0[:, :5] * 2.0)   # rank-major order
        assert r[2].shape == (64,)
        assert r[3]["seconds_max"] == 0.2 and r[3]["converged_total"] == int((res[0][2] == 0).sum())


def _worker_uneven(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mobile_manipulator_mpc_b200 import sharding
    B = 33                                                   # strong scaling, B % world != 0: shards of 17 and 16
    lo, hi = sharding.shard_bounds(B, world, rank)
    idx = torch.arange(lo, hi, dtype=torch.float64)
    u0 = idx[:, None] * torch.ones(1, 5, dtype=torch.float64)
    status = torch.arange(lo, hi, dtype=torch.int32) % 3
    allu0, allst = sharding.gather_results(dist, u0, status, B=B)
    q.put((rank, allu0.numpy(), allst.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_of_uneven_shards():
    world, port = 2, 29613
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_uneven, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in ps]
    for _, u, st in res:
        assert u.shape == (33, 5) and np.array_equal(u[:, 0], np.arange(33.0)) and np.array_equal(st, np.arange(33) % 3)
