"""Multi-GPU plumbing for the risk evaluations: shard independent trees, all-reduce 3 doubles.

The path shards on independent units (trees; for CLIP the matched pair index, whose K+1 text and
K+1 image trees stay on one rank).  No data-path collective exists: the only exchange is one
all-reduce of {sum, sum of squares, count} per risk evaluation (NCCL on GPUs; gloo in the CPU
tests).  Philox counters are keyed by the GLOBAL tree index, so any sharding draws the same trees.
"""
import torch


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of ``n`` units for ``rank`` of ``world`` (sizes differ by at most 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def dist_info(group=None):
    """(rank, world) of the default/initialised process group, (0, 1) when not distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_reduce_sums(sums, group=None):
    """In-place SUM all-reduce of the float64[3] risk accumulator across ranks (no-op when world == 1)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def mean_se_from_sums(sums, se_count=None):
    """(mean, population-std / sqrt(se_count or count)) from {sum, sumsq, count}."""
    s1, s2, c = (float(x) for x in torch.as_tensor(sums).tolist())
    if c <= 0:
        raise ValueError("empty risk accumulator")
    mean = s1 / c
    var = max(s2 / c - mean * mean, 0.0)
    return mean, (var ** 0.5) / (float(se_count if se_count else c) ** 0.5)
