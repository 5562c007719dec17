"""One double-precision box (s,mu) count through the pre-filter (for ncu / the diagnostics build):
python tools/prof_pf.py [lib] [N] [L]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200.api as api
from pathlib import Path
if len(sys.argv) > 1 and sys.argv[1] != "-":
    api.LIB_PATH = Path(sys.argv[1]).resolve()
import fcfc_b200 as F
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
L = float(sys.argv[3]) if len(sys.argv) > 3 else 1169.6
F.init(devices=[0])
rng = np.random.default_rng(20261017)
x = rng.random((N, 3)) * L
b = F.Bins(periodic=True, prec="double", arith=1, box=L, bintype=1, smax=200., ds=5., nmu=120)
g = F.Catalog(x[:, 0], x[:, 1], x[:, 2], bins=b)
for _ in range(2):
    c = F.count_pairs(g, None, b); st = F.stats()
    print(f"kernel {st['ms_count']:.2f} ms prefilter={st['prefilter']} evals {st['pair_evals']:.4g} pairs {int(c.sum())}", flush=True)
