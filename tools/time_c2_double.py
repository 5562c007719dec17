"""Kernel time of the bench workload in double precision for one or more builds: python tools/time_c2_double.py lib [lib ...]"""
import sys, os, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
import fcfc_b200.api as api
rng = np.random.default_rng(20261017)
N, L = 10_000_000, 2000.0
x = [np.ascontiguousarray(rng.random(N) * L) for _ in range(3)]
for lib in sys.argv[1:]:
    api.LIB_PATH = type(api.LIB_PATH)(lib); api._lib = None
    F.init(devices=[0])
    for bt in (1, 0):
        b = F.Bins(periodic=True, prec="double", arith=1, box=L, bintype=bt, smax=200., ds=5., nmu=120)
        g = F.Catalog(*x, bins=b)
        best = 1e30
        for _ in range(3):
            c = F.count_pairs(g, None, b); st = F.stats(); best = min(best, st["ms_count"])
        print(f"{lib.split('/')[-2]} bt={bt}: kernel {best:.1f} ms, kind {st['prefilter']}, pairs {int(c.sum())} digest {hashlib.sha1(c.tobytes()).hexdigest()[:12]}", flush=True)
        g.destroy()
