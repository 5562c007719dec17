"""Scale check at 10^8 points (config C4 size, reduced s_max): python tools/time_big.py [N] [L] [smax]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
L = float(sys.argv[2]) if len(sys.argv) > 2 else 2000.0
smax = float(sys.argv[3]) if len(sys.argv) > 3 else 50.0
F.init()
rng = np.random.default_rng(7)
t0 = time.time()
x = np.empty((3, N), dtype=np.float32)
for d in range(3):
    x[d] = rng.random(N, dtype=np.float32) * np.float32(L)
x[x >= L] = np.nextafter(np.float32(L), np.float32(0))
print(f"generated {N} points in {time.time() - t0:.1f} s", flush=True)
tot = {}
for bt in (1, 0):
    b = F.Bins(periodic=True, prec="float", arith=1, box=L, bintype=bt, smax=smax, ds=5., nmu=120)
    t0 = time.time()
    g = F.Catalog(x[0], x[1], x[2], bins=b)
    t1 = time.time()
    c = F.count_pairs(g, None, b); st = F.stats()
    t2 = time.time()
    tot[bt] = int(c.sum())
    exp = 0.5 * N * (N - 1) * (4 / 3 * np.pi * smax ** 3) / L ** 3
    print(f"bt={bt}: upload {t1 - t0:.2f} s, count call {t2 - t1:.2f} s (kernel {st['ms_count']:.1f} ms, sort {st['ms_sort']:.1f} ms), "
          f"pairs {tot[bt]} (expected {exp:.6g}, ratio {tot[bt] / exp:.6f}), evals {st['pair_evals']:.4g}, grid {st['ncell']}, items {st['nitem']}", flush=True)
    if bt == 0:
        parts = [F.count_pairs(g, None, b, part=p, nparts=3) for p in range(3)]
        print("shards add up:", bool((sum(parts) == c).all()))
    g.destroy()
print("ISO total - SMU total (pairs with mu rounding to 1, dropped as in the reference):", tot[0] - tot[1])
