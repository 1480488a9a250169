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


def gather_results(dist, u0, status, B=None):
    """all_gather of the applied controls u0 [b,5] and the per-instance status, rank-major.  With ``B`` (the global batch
    size) the shards may be uneven, as shard_bounds produces them when B % world != 0: every rank pads its shard to the
    common size ceil(B / world) for the collective and the padding is cut out of the result."""
    world = dist.get_world_size()
    per = u0.shape[0] if B is None else (B + world - 1) // world
    if u0.shape[0] < per:   # the last rank(s) of an uneven split
        pad = per - u0.shape[0]
        u0 = torch.cat([u0, u0.new_zeros((pad,) + tuple(u0.shape[1:]))])
        status = torch.cat([status, status.new_full((pad,), -1)])
    allu0 = torch.empty((world * per,) + tuple(u0.shape[1:]), dtype=u0.dtype, device=u0.device)
    allst = torch.empty((world * per,), dtype=status.dtype, device=status.device)
    dist.all_gather_into_tensor(allu0, u0.contiguous())
    dist.all_gather_into_tensor(allst, status.contiguous())
    if B is not None and world * per != B:
        keep = torch.cat([torch.arange(r * per, r * per + (shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0]))
                          for r in range(world)]).to(allu0.device)
        allu0, allst = allu0.index_select(0, keep), allst.index_select(0, keep)
    return allu0, allst


def reduce_stats(dist, converged, seconds, device="cpu"):
    """whole-job numbers: converged instances summed, elapsed time maxed over ranks."""
    c = torch.tensor([float(converged)], dtype=torch.float64, device=device)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return dict(converged_total=int(c.item()), seconds_max=float(t.item()))


def run_interleaved(shards, step, steps):
    """Drive K independent sub-batches of one GPU concurrently: one host thread and one CUDA stream per shard, so that the
    thin tail of one shard's solve (a few slow instances, launch-latency bound) overlaps the bulk of another's.  ``shards``:
    objects with their own solver handles (handles are per host thread); ``step(shard)`` advances one of them by one MPC step
    and may return a value; returns the list of per-shard lists of those values.  Used by the closed-loop and episode
    benches; a single solve call cannot hide its own tail, independent batches can."""
    import threading
    import torch
    main = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in shards]
    for s in streams:
        s.wait_stream(main)
    out = [[] for _ in shards]
    err = []

    def work(i):
        try:
            with torch.cuda.stream(streams[i]):
                for _ in range(steps):
                    out[i].append(step(shards[i]))
        except BaseException as e:  # surfaced by the caller
            err.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(shards))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for s in streams:
        main.wait_stream(s)
    if err:
        raise err[0]
    return out
