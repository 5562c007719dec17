"""CPU: the N>1 plumbing of fcfc_b200/sharding.py (what bench.py runs under torchrun) with world_size 2 over gloo:
the sliced upload + all-gather rebuilds every catalogue column bit for bit, the engine's strided deal of the work items
covers every item exactly once, and the all-reduced per-rank histograms equal the full count (partial counts come from
the oracle here, dealt the way the engine deals its items; on GPUs they come from fcfc_gpu_count_partial:
test_gpu_parity.py::test_full_size_properties_c1, test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import box_catalog
from fcfc_b200.sharding import allgather_columns, allreduce_histogram, shard_items, slice_bounds, slice_len
from oracle import oracle


def test_strided_deal_and_slices_partition_exactly():
    for nitem in (0, 1, 7, 100, 117649):
        for nparts in (1, 2, 3, 8):
            seen = np.zeros(nitem, dtype=int)
            for p in range(nparts):
                seen[list(shard_items(nitem, p, nparts))] += 1
            assert (seen == 1).all()
            # cost-sorted order dealt round-robin: shard sizes differ by at most one item
            sizes = [len(shard_items(nitem, p, nparts)) for p in range(nparts)]
            assert max(sizes) - min(sizes) <= 1
            rows = []
            for p in range(nparts):
                b, e = slice_bounds(nitem, p, nparts)
                assert 0 <= b <= e <= nitem and e - b <= slice_len(nitem, nparts)
                rows += list(range(b, e))
            assert rows == list(range(nitem))
    with pytest.raises(ValueError):
        shard_items(10, 2, 2)
    with pytest.raises(ValueError):
        slice_bounds(10, -1, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = dict(box=300.0, bintype=1, smax=30.0, ds=1.5, nmu=20)
    D = box_catalog(3001, 300.0, 81, weights=False)        # odd size: the last slice is padded
    R = box_catalog(2000, 300.0, 82, weights=False)
    # data: every rank holds only its slice of D, the all-gather rebuilds the columns
    b, e = slice_bounds(len(D[0]), rank, world)
    cols = allgather_columns([torch.from_numpy(np.ascontiguousarray(c[b:e])) for c in D], len(D[0]))
    same = all(np.array_equal(g.numpy(), c) for g, c in zip(cols, D))
    Dg = tuple(g.numpy() for g in cols)
    ob = oracle.setup(prec="d", periodic=True, **kw)
    pd, pr = oracle.preprocess(ob, Dg), oracle.preprocess(ob, R)
    # work: primaries dealt round-robin like the engine's work items, secondary replicated
    mine = np.array(list(shard_items(len(Dg[0]), rank, world)))
    part = {k: (v[mine] if v is not None else None) for k, v in pd.items()}
    h = torch.from_numpy(oracle.count(ob, part, pr))
    allreduce_histogram(h)
    if rank == 0:
        full = oracle.count(ob, pd, pr)
        out.put(same and bool(np.array_equal(h.numpy(), full)) and int(full.sum()) > 0)
    dist.destroy_process_group()


def test_two_rank_allgather_and_allreduce_equal_full_count():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
