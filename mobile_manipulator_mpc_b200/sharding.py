"""Multi-GPU host logic: the batch shards trivially (independent instances), one process per GPU,
no collective inside the solve.  NCCL (gloo in the CPU tests) only gathers per-instance results
and reduces statistics -- SURVEY.md 8(e)."""
import torch


def rank_seed(base_seed, rank):
    """Every rank draws its own shard of the synthetic workload (weak scaling)."""
    return int(base_seed) + 1000 * int(rank)


def shard_bounds(B, world, rank):
    """Contiguous slice [lo, hi) of a global batch of B instances owned by ``rank`` (strong scaling)."""
    per = (B + world - 1) // world
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def gather_results(dist, u0, status):
    """all_gather of the applied controls u0 [B,5] and the per-instance status, rank-major."""
    world = dist.get_world_size()
    allu0 = torch.empty((world * u0.shape[0],) + tuple(u0.shape[1:]), dtype=u0.dtype, device=u0.device)
    allst = torch.empty((world * status.shape[0],), dtype=status.dtype, device=status.device)
    dist.all_gather_into_tensor(allu0, u0.contiguous())
    dist.all_gather_into_tensor(allst, status.contiguous())
    return allu0, allst


def reduce_stats(dist, converged, seconds, device="cpu"):
    """whole-job numbers: converged instances summed, elapsed time maxed over ranks."""
    c = torch.tensor([float(converged)], dtype=torch.float64, device=device)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return dict(converged_total=int(c.item()), seconds_max=float(t.item()))
