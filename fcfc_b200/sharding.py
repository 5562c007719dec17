"""Host-side logic of the multi-GPU decomposition.

The primary catalogue's work items are split into `nparts` contiguous ranges (fcfc_gpu_count_partial does the same
split on the device side, engine.cu: item_begin/item_end), the secondary catalogue is replicated, and the per-rank
histograms are summed with one all-reduce.  Integer histograms are exact for any number of ranks."""
from __future__ import annotations


def item_range(nitem: int, part: int, nparts: int) -> tuple[int, int]:
    """[begin, end) of a contiguous split of `nitem` units into `nparts` (used to shard host-side arrays)."""
    if not (0 <= part < nparts):
        raise ValueError("invalid shard")
    return nitem * part // nparts, nitem * (part + 1) // nparts


def shard_items(nitem: int, part: int, nparts: int) -> range:
    """Work items of shard `part` as the engine assigns them (engine.cu / count_kernel.cuh): the items are sorted
    by decreasing estimated cost and shard `part` takes positions part, part + nparts, ... of that order."""
    if not (0 <= part < nparts):
        raise ValueError("invalid shard")
    return range(part, nitem, nparts)


def allreduce_histogram(hist, group=None):
    """Sum a per-rank histogram (torch tensor, int64 or float64, on the rank's device) over all ranks in place.
    NCCL on GPUs (NVLink/NVSwitch), gloo in the CPU tests."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist
