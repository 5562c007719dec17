import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fcfc_b200 as F
from cases import clustered_box_catalog, box_catalog
F.init()
N, L = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000, float(sys.argv[2]) if len(sys.argv) > 2 else 1169.6
for name, cat in (("uniform", box_catalog(N, L, 3, weights=False)), ("clustered", clustered_box_catalog(N, L, 4)[:3])):
    b = F.Bins(periodic=True, prec="float", arith=1, box=L, bintype=1, smax=200., ds=5., nmu=120)
    g = F.Catalog(*cat, bins=b)
    F.count_pairs(g, None, b); c = F.count_pairs(g, None, b); st = F.stats()
    parts = [F.count_pairs(g, None, b, part=p, nparts=4) for p in range(4)]
    pt = []
    for p in range(4):
        F.count_pairs(g, None, b, part=p, nparts=4); pt.append(F.stats()["ms_count"])
    ok = np.array_equal(sum(parts), c)
    print(f"{name}: kernel {st['ms_count']:.1f} ms, pairs {int(c.sum()):.4g}, evals {st['pair_evals']:.4g}, grid {st['ncell']}, items {st['nitem']}; 4 shards {[round(x,1) for x in pt]} ms sum_ok={ok}", flush=True)
    g.destroy()
