import sys, os, time, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
import fcfc_b200.api as api
libs = sys.argv[1:]
rng = np.random.default_rng(20261017)
N, L = 10_000_000, 2000.0
x = (rng.random((N, 3)) * L).astype(np.float32)
for lib in libs:
    api.LIB_PATH = type(api.LIB_PATH)(lib); api._lib = None
    F.init()
    for bt in (1, 0):
        b = F.Bins(periodic=True, prec="float", arith=1, box=L, bintype=bt, smax=200., ds=5., nmu=120)
        g = F.Catalog(x[:, 0], x[:, 1], x[:, 2], bins=b)
        for _ in range(2):
            c = F.count_pairs(g, None, b); st = F.stats()
        print(f"{os.path.basename(lib)} bt={bt}: kernel {st['ms_count']:.1f} ms, pairs {int(c.sum())}, evals {st['pair_evals']:.4g}", flush=True)
        g.destroy()
