"""Cell-size scan at the density of configs[3] at size (10^8 points in L = 3000): 1.25x10^7 uniform points in L = 1500, float (s,mu).
python tools/time_density.py [k ...]   (0 = the engine's own choice)"""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
F.init(devices=[0])
ks = [int(v) for v in sys.argv[1:]] or [0, 5, 6, 7, 8]
rng = np.random.default_rng(20261017)
N, L = 12_500_000, 1500.0
x = [np.ascontiguousarray(rng.random(N) * L) for _ in range(3)]
b = F.Bins(periodic=True, prec="float", arith=1, box=L, bintype=1, smax=200., ds=5., nmu=120)
g = F.Catalog(*x, bins=b)
for k in ks:
    F.set_option("defaults", 0)
    if k:
        F.set_option("k", k)
    best = 1e30
    for _ in range(2):
        c = F.count_pairs(g, None, b); st = F.stats(); best = min(best, st["ms_count"])
    print(f"k={k}: kernel {best:.1f} ms evals {st['pair_evals']:.4g} computed {st['pair_evals_computed']:.4g} grid {st['ncell']} items {st['nitem']} digest {hashlib.sha1(c.tobytes()).hexdigest()[:12]}", flush=True)
