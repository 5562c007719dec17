import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
F.init()
rng = np.random.default_rng(5)
N, L = 2_000_000, 1169.6
x = rng.random((N, 3)) * L; w = rng.uniform(0.75, 1.25, N)
for prec in ("float", "double"):
    for bt in (0, 1):
        b = F.Bins(periodic=True, prec=prec, arith=1, box=L, bintype=bt, smax=200., ds=5., nmu=120)
        g = F.Catalog(x[:, 0], x[:, 1], x[:, 2], w, bins=b)
        for wt in (False, True):
            F.count_pairs(g, None, b, withwt=wt); c = F.count_pairs(g, None, b, withwt=wt); st = F.stats()
            print(f"{prec} bt={bt} weighted={wt}: {st['ms_count']:.1f} ms sum={c.sum():.8g}", flush=True)
        g.destroy()
