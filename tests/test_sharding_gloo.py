"""CPU: the N>1 plumbing with world_size 2 over gloo -- shard ranges partition the work exactly once, and the
all-reduced per-rank histograms equal the full count (partial counts come from the oracle here; on GPUs they come
from fcfc_gpu_count_partial, tested in test_gpu_parity.py::test_full_size_properties_c1)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import box_catalog
from fcfc_b200.sharding import allreduce_histogram, item_range, shard_items
from oracle import oracle


def test_item_ranges_partition_exactly():
    for nitem in (0, 1, 7, 100, 117649):
        for nparts in (1, 2, 3, 8):
            seen = []
            for p in range(nparts):
                b, e = item_range(nitem, p, nparts)
                assert 0 <= b <= e <= nitem
                seen += list(range(b, e)) if nitem < 1000 else [(b, e)]
            if nitem < 1000:
                assert seen == list(range(nitem))
            else:
                assert seen[0][0] == 0 and seen[-1][1] == nitem and all(seen[i][1] == seen[i + 1][0] for i in range(nparts - 1))
    with pytest.raises(ValueError):
        item_range(10, 2, 2)
    for nitem in (0, 1, 7, 100):
        for nparts in (1, 2, 3, 8):
            seen = sorted(i for p in range(nparts) for i in shard_items(nitem, p, nparts))
            assert seen == list(range(nitem))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = dict(box=300.0, bintype=1, smax=30.0, ds=1.5, nmu=20)
    D = box_catalog(3000, 300.0, 81, weights=False)
    R = box_catalog(2000, 300.0, 82, weights=False)
    ob = oracle.setup(prec="d", periodic=True, **kw)
    pd, pr = oracle.preprocess(ob, D), oracle.preprocess(ob, R)
    b, e = item_range(len(D[0]), rank, world)          # shard the primaries, replicate the secondary
    part = {k: (v[b:e] if v is not None else None) for k, v in pd.items()}
    h = torch.from_numpy(oracle.count(ob, part, pr))
    allreduce_histogram(h)
    if rank == 0:
        full = oracle.count(ob, pd, pr)
        out.put(bool(np.array_equal(h.numpy(), full)) and int(full.sum()) > 0)
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_full_count():
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
