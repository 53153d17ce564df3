"""Batch sharding across GPUs: polynomials are independent (the reference API is one polynomial
per call, src/unordered.rs:826), so ranks take contiguous row ranges and never communicate on
the data path.  torch.distributed is used only for the barrier and the max-over-ranks timing."""


def shard_rows(batch, world_size, rank):
    """Contiguous split of `batch` rows: returns (first_row, row_count) for `rank`.
    Ranks differ by at most one row; every row is owned by exactly one rank."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    base, extra = divmod(batch, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def max_over_ranks(value, dist=None, device=None):
    """Whole-job time = max over ranks (all-reduce MAX); identity without a process group."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank, seconds_per_rank, dist=None, device=None):
    """units all ranks processed / max-over-ranks time (bench.py's `value`)."""
    if dist is None or not dist.is_initialized():
        return units_per_rank / seconds_per_rank
    import torch

    u = torch.tensor([float(units_per_rank)], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / max_over_ranks(seconds_per_rank, dist, device)
