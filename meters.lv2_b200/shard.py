"""Multi-GPU plumbing: channel-instance sharding (no data-path collective) and the one optional exchange.

Every metering instance is self-contained, so N GPUs simply own contiguous instance ranges:
rank r of W processes instances  [r*n/W, (r+1)*n/W).  The only cross-instance quantity is the whole-mix gated
loudness: the int32 sum of all instances' hist_M / hist_S (+ counts), reduced per GPU by
b200m_ebu_mix_reduce and across GPUs by ONE all-reduce of B200M_MIX_WORDS int32 (NCCL on GPUs, gloo in CPU
tests); integer sums are order independent, hence bit-exact for any GPU count.
"""
MIX_WORDS = 1508


def shard_range(n_total, rank, world):
    """contiguous, balanced partition: returns (first, count) of the instances owned by `rank`."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad rank/world/n_total")
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi - lo


def shard_rows(n_total, nchan, rank, world):
    """row range of a planar [n_total*nchan, nfram] batch owned by `rank`."""
    lo, cnt = shard_range(n_total, rank, world)
    return lo * nchan, cnt * nchan


def allreduce_mix(mix, group=None):
    """sum the per-rank whole-mix histogram vector (int32[MIX_WORDS]) over all ranks, in place."""
    import torch
    import torch.distributed as dist
    if mix.dtype != torch.int32 or mix.numel() != MIX_WORDS:
        raise ValueError("mix must be int32[%d]" % MIX_WORDS)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mix, op=dist.ReduceOp.SUM, group=group)
    return mix


def gather_results(local, group=None):
    """concatenate per-rank result arrays (numpy structured or plain) on every rank, in rank order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out
