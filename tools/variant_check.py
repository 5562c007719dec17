"""Time one build of libfcfc_b200.so on a fixed set of workloads and print, per workload, the kernel time and a
digest of the counts (run under gpurun; tools/pick_variant.py compares the outputs of several builds).

    python tools/variant_check.py <path/to/libfcfc_b200.so> [quick]
"""
import sys, os, json, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200.api as api
from pathlib import Path
api.LIB_PATH = Path(sys.argv[1]).resolve()
import fcfc_b200 as F
quick = len(sys.argv) > 2
F.init(devices=[0])
rng = np.random.default_rng(20261017)
out = {"lib": sys.argv[1]}


def digest(c):
    return hashlib.sha1(np.ascontiguousarray(c).tobytes()).hexdigest()[:16]


def run(name, cats, bins, pairs, withwt=False, reps=2):
    g = [F.Catalog(*c, bins=bins) for c in cats]
    for p in pairs:
        i, j = p
        best = 1e30
        for _ in range(reps):
            c = F.count_pairs(g[i], None if i == j else g[j], bins, withwt=withwt)
            st = F.stats()
            best = min(best, st["ms_count"])
        key = f"{name}:{i}{j}"
        out[key] = {"ms": round(best, 3), "sum": float(c.sum()), "digest": digest(c) if not withwt else None,
                    "wsums": [float(v) for v in c[:: max(1, len(c) // 16)]] if withwt else None,
                    "evals": st["pair_evals"]}
        print(key, out[key], flush=True)
    for h in g:
        h.destroy()


def box(n, L):
    x = rng.random((n, 3)) * L
    return (x[:, 0], x[:, 1], x[:, 2])


def survey(n):
    ra = np.deg2rad(rng.uniform(120, 180, n)); sd = rng.uniform(0, 0.5, n); cd = np.sqrt(1 - sd * sd)
    d = rng.uniform(1000, 1700, n)
    return (d * cd * np.cos(ra), d * cd * np.sin(ra), d * sd, rng.uniform(0.75, 1.25, n))


kw = dict(periodic=True, box=2000.0, smax=200.0, ds=5.0)
c7 = box(10_000_000 if not quick else 2_000_000, 2000.0 if not quick else 1169.6)
kw["box"] = 2000.0 if not quick else 1169.6
run("c2_f_smu_fma", [c7], F.Bins(prec="float", arith=1, bintype=1, nmu=120, **kw), [(0, 0)], reps=3)
run("c2_f_iso_fma", [c7], F.Bins(prec="float", arith=1, bintype=0, **kw), [(0, 0)])
del c7
kw["box"] = 1169.6
c6 = box(2_000_000, 1169.6)
c6b = box(1_000_000, 1169.6)
run("2e6_f_smu_scalar", [c6, c6b], F.Bins(prec="float", arith=0, bintype=1, nmu=120, **kw), [(0, 0), (0, 1)])
run("2e6_d_smu_fma", [c6], F.Bins(prec="double", arith=1, bintype=1, nmu=120, **kw), [(0, 0)])
run("2e6_d_iso_scalar", [c6], F.Bins(prec="double", arith=0, bintype=0, **kw), [(0, 0)])
run("2e6_f_smu_nmu37", [c6], F.Bins(prec="float", arith=1, bintype=1, nmu=37, periodic=True, box=1169.6, smax=150.0, ds=2.0), [(0, 0)])
w = rng.uniform(0.75, 1.25, 2_000_000)
run("2e6_f_smu_wt", [c6 + (w,)], F.Bins(prec="float", arith=1, bintype=1, nmu=120, **kw), [(0, 0)], withwt=True)
D, R = survey(200_000), survey(1_000_000)
skw = dict(periodic=False, bintype=1, smax=200.0, ds=5.0, nmu=100)
run("svy_f_smu", [D[:3], R[:3]], F.Bins(prec="float", arith=1, **skw), [(0, 0), (0, 1)])
run("svy_d_smu", [D[:3], R[:3]], F.Bins(prec="double", arith=0, **skw), [(0, 0), (0, 1)])
print("JSON " + json.dumps(out))
