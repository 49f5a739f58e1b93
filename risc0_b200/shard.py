"""Segment sharding across GPUs: continuation segments are independent units, so segment `s` goes to
rank `s mod world` and the data path has NO collective. torch.distributed is used only for the barrier and for the
max-over-ranks time / result exchange of the benchmark (NCCL on GPUs, gloo in the CPU tests). The reference runs one
prover process per GPU the same way (risc0/r0vm/src/actors/mod.rs:449-456, worker.rs:70-76)."""


def assign_segments(num_segments, world, rank):
    """indices of the segments rank `rank` proves (static round-robin)"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank: %r/%r" % (world, rank))
    return list(range(rank, num_segments, world))


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity when torch.distributed is not initialised)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(value, device=None):
    """every rank's integer (e.g. segments proved), as a list indexed by rank"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(value)]
    t = torch.tensor([int(value)], dtype=torch.int64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]


def whole_job_throughput(units_per_rank, seconds_max):
    """value = units all ranks processed / max-over-ranks time"""
    return sum(units_per_rank) / seconds_max
