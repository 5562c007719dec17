"""Host-side plumbing of the multi-GPU decomposition, one process per GPU.

* Work: the engine sorts the primary catalogue's work items by decreasing estimated cost and shard `part` of `nparts`
  takes positions part, part + nparts, ... of that order (fcfc_gpu_count_partial; count_kernel.cuh, persistent warp
  loop).  `shard_items` states the same rule for host-side callers and tests.
* Data: the secondary catalogue is replicated.  Instead of every rank uploading the whole catalogue over PCIe, rank r
  uploads the contiguous slice `slice_bounds(n, r, world)` of every column and the ranks all-gather the slices over
  NVLink (`allgather_columns`: NCCL on GPUs, gloo in the CPU tests) -- the replacement of the reference's
  kdtree_broadcast (src/tree/kdtree.c:529-619).
* Result: one all-reduce of the per-rank histograms (`allreduce_histogram`), replacing MPI_Ireduce
  (src/fcfc/2pt_box/count_func.c:7654-7721).  Integer histograms are exact for any number of ranks.
"""
from __future__ import annotations


def shard_items(nitem: int, part: int, nparts: int) -> range:
    """Positions of the cost-sorted work-item order that shard `part` processes."""
    if not (0 <= part < nparts):
        raise ValueError("invalid shard")
    return range(part, nitem, nparts)


def slice_len(n: int, world: int) -> int:
    """Rows per rank of the sliced upload (the last slices are padded: all-gather needs equal sizes)."""
    return (n + world - 1) // world


def slice_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of the rows of a catalogue column that rank `rank` uploads."""
    if not (0 <= rank < world):
        raise ValueError("invalid rank")
    m = slice_len(n, world)
    return min(n, rank * m), min(n, (rank + 1) * m)


def allgather_columns(local_cols, n: int, group=None):
    """local_cols: this rank's slices (1-D torch tensors on the rank's device, one per catalogue column, rows
    slice_bounds(n, rank, world)).  Returns the full columns (length n) on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return [c[:n] for c in local_cols]
    m = slice_len(n, world)
    out = []
    for c in local_cols:
        if c.numel() < m:                       # pad the short last slice(s)
            c = torch.cat([c, c.new_zeros(m - c.numel())])
        full = c.new_empty(m * world)
        dist.all_gather_into_tensor(full, c.contiguous(), group=group)
        out.append(full[:n])
    return out


def allreduce_histogram(hist, group=None):
    """Sum a per-rank histogram (torch tensor, int64 or float64, on the rank's device) over all ranks in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist
