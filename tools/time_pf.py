"""Double precision with and without the float pre-filter (engine option no_prefilter): kernel times and equality of
the counts (unweighted: bit for bit; weighted: 1e-12) on the C2 box, a 2x10^6 box in three binnings, and survey counts."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
F.init(devices=[0])
rng = np.random.default_rng(20261017)
quick = len(sys.argv) > 1


def survey(n, seed):
    r = np.random.default_rng(seed)
    ra = np.deg2rad(r.uniform(120, 180, n)); sd = r.uniform(0, 0.5, n); cd = np.sqrt(1 - sd * sd)
    d = np.cbrt(r.uniform(1000.0 ** 3, 1700.0 ** 3, n))
    return (d * cd * np.cos(ra), d * cd * np.sin(ra), d * sd, r.uniform(0.75, 1.25, n))


MODES = (("plain", dict(no_df=1, no_prefilter=1)), ("prefilter", dict(no_df=1, force_prefilter=1)), ("auto", dict()))
KIND = {0: "plain FP64", 1: "pre-filter", 2: "float-speed"}


def run(tag, bins, cats, pairs, withwt):
    g = [F.Catalog(*(c if withwt else c[:3]), bins=bins) for c in cats]
    for (i, j) in pairs:
        res = {}
        for name, opts in MODES:
            F.set_option("defaults", 0)
            for k, v in opts.items():
                F.set_option(k, v)
            best = 1e30
            for _ in range(2):
                c = F.count_pairs(g[i], None if i == j else g[j], bins, withwt=withwt); st = F.stats()
                best = min(best, st["ms_count"])
            res[name] = (best, c, st["prefilter"], st["pair_evals"])
        F.set_option("defaults", 0)
        ref = res["plain"][1]
        same = all(bool(np.array_equal(ref, r[1])) if not withwt else bool(np.allclose(ref, r[1], rtol=1e-12, atol=0) and np.array_equal(ref == 0, r[1] == 0))
                   for r in res.values())
        print(f"{tag} pair {i}{j} wt={int(withwt)}: plain {res['plain'][0]:.1f} ms | pre-filter {res['prefilter'][0]:.1f} ms (used={res['prefilter'][2]}) | "
              f"auto [{KIND[res['auto'][2]]}] {res['auto'][0]:.1f} ms x{res['plain'][0] / res['auto'][0]:.2f} | same={same} sum={float(ref.sum()):.6g} evals={res['auto'][3]:.4g}", flush=True)
    for h in g:
        h.destroy()


n7, L7 = (10_000_000, 2000.0) if not quick else (2_000_000, 1169.6)
x = rng.random((n7, 3)) * L7
c7 = (x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy())
del x
for arith in (1, 0):
    run(f"box smu {n7:.0e} arith={arith}", F.Bins(periodic=True, prec="double", arith=arith, box=L7, bintype=1, smax=200., ds=5., nmu=120), [c7], [(0, 0)], False)
run(f"box iso {n7:.0e}", F.Bins(periodic=True, prec="double", arith=1, box=L7, bintype=0, smax=200., ds=5.), [c7], [(0, 0)], False)
del c7
x = rng.random((2_000_000, 3)) * 1169.6
c6 = (x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy(), rng.uniform(0.75, 1.25, 2_000_000))
run("box spi 2e6", F.Bins(periodic=True, prec="double", arith=1, box=1169.6, bintype=2, smax=100., ds=5., pmin=0., pmax=120., dpi=4.), [c6], [(0, 0)], False)
run("box smu 2e6 weighted", F.Bins(periodic=True, prec="double", arith=1, box=1169.6, bintype=1, smax=200., ds=5., nmu=120), [c6], [(0, 0)], True)
run("box iso 2e6 weighted", F.Bins(periodic=True, prec="double", arith=0, box=1169.6, bintype=0, smax=200., ds=5.), [c6], [(0, 0)], True)
run("box spi 2e6 weighted", F.Bins(periodic=True, prec="double", arith=1, box=1169.6, bintype=2, smax=100., ds=5., pmin=0., pmax=120., dpi=4.), [c6], [(0, 0)], True)
D, R = survey(200_000, 1), survey(2_000_000, 2)
run("survey spi weighted", F.Bins(periodic=False, prec="double", arith=0, bintype=2, smax=40., ds=2., pmin=0., pmax=80., dpi=1.), [D, R], [(0, 0), (0, 1), (1, 1)], True)
run("survey spi", F.Bins(periodic=False, prec="double", arith=1, bintype=2, smax=40., ds=2., pmin=0., pmax=80., dpi=1.), [D, R], [(0, 1)], False)
run("survey smu", F.Bins(periodic=False, prec="double", arith=1, bintype=1, smax=200., ds=5., nmu=100), [D, R], [(0, 1)], False)
run("survey iso", F.Bins(periodic=False, prec="double", arith=0, bintype=0, smax=200., ds=5.), [D, R], [(0, 1)], False)
