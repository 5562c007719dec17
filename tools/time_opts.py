"""A/B of engine options on the bench workloads: python tools/time_opts.py "inplace=0" "inplace=1" ...
Each argument is a comma-separated option set; prints kernel times (best of 3) for C2 (s,mu) and isotropic, float FMA order,
and a digest of the counts (all sets must agree)."""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
F.init(devices=[0])
sets = sys.argv[1:] or ["inplace=0", "inplace=1"]
rng = np.random.default_rng(20261017)
N, L = 10_000_000, 2000.0
x = [np.ascontiguousarray(rng.random(N) * L) for _ in range(3)]
for bt in (1, 0):
    b = F.Bins(periodic=True, prec="float", arith=1, box=L, bintype=bt, smax=200., ds=5., nmu=120)
    g = F.Catalog(*x, bins=b)
    for s in sets:
        F.set_option("defaults", 0)
        for kv in s.split(","):
            if kv:
                k, v = kv.split("=")
                F.set_option(k, int(v))
        best = 1e30
        for _ in range(3):
            c = F.count_pairs(g, None, b); st = F.stats()
            best = min(best, st["ms_count"])
        print(f"bt={bt} [{s}]: kernel {best:.1f} ms evals {st['pair_evals']:.4g} computed {st['pair_evals_computed']:.4g} pairs {int(c.sum())} digest {hashlib.sha1(c.tobytes()).hexdigest()[:12]} grid {st['ncell']}", flush=True)
    F.set_option("defaults", 0)
    g.destroy()
