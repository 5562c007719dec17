"""Run one workload a few times (for ncu): python tools/prof_one.py N L bintype prec arith [reps]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
N, L, bt, prec, arith = int(float(sys.argv[1])), float(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
F.init()
rng = np.random.default_rng(20261017)
x = rng.random((N, 3)) * L
b = F.Bins(periodic=True, prec=prec, arith=arith, box=L, bintype=bt, smax=200., ds=5., nmu=120)
g = F.Catalog(x[:, 0], x[:, 1], x[:, 2], bins=b)
for _ in range(reps):
    c = F.count_pairs(g, None, b); st = F.stats()
    print(f"kernel {st['ms_count']:.2f} ms evals {st['pair_evals']:.4g} pairs {st['pairs_in']:.5g} -> {st['pair_evals']/st['ms_count']*1e-9:.3f} Tevals/s grid {st['ncell']} items {st['nitem']}", flush=True)
