"""What one rank of an N-GPU run sees, measured on one GPU: the time of shard 0 of N (fcfc_gpu_count_partial) against
1/N of the full count, for several work-item granularities (engine option items_per_warp)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
F.init(devices=[0])
rng = np.random.default_rng(20261017)
N, L = 10_000_000, 2000.0
x = (rng.random((N, 3)) * L).astype(np.float32)
b = F.Bins(periodic=True, prec="float", arith=1, box=L, bintype=1, smax=200., ds=5., nmu=120)
g = F.Catalog(x[:, 0], x[:, 1], x[:, 2], bins=b)
for ipw in (8, 16, 32, 64):
    F.set_option("items_per_warp", ipw)
    row = []
    for nparts in (1, 4, 8):
        best = 1e30
        for _ in range(2):
            F.count_pairs(g, None, b, part=0, nparts=nparts); st = F.stats()
            best = min(best, st["ms_count"])
        row.append(best)
    print(f"items_per_warp={ipw}: full {row[0]:.1f} ms | shard 0/4 {row[1]:.1f} ms (eff {row[0] / 4 / row[1]:.3f}) | shard 0/8 {row[2]:.1f} ms (eff {row[0] / 8 / row[2]:.3f}) items {st['nitem']}", flush=True)
