"""Small counts of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
float/double x fast/generic x box (s,mu), isotropic, (s_perp,pi) and survey, weighted and unweighted, with shallow stacks
so that the drains, the flagged-pair corrections and the dense-cell path all run; each result is checked against the
oracle (the checker, not the product)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fcfc_b200 as F
from oracle import oracle
from cases import box_catalog, survey_catalog

F.init(devices=[0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
cat = box_catalog(n, 120.0, 5)
D, R = survey_catalog(n // 2, 6), survey_catalog(n, 7)
bad = 0


def check(tag, periodic, prec, cats, kw, withwt, arith, opts=()):
    global bad
    for k, v in opts:
        F.set_option(k, v)
    b = F.Bins(periodic=periodic, prec=prec, arith=arith, **kw)
    g = [F.Catalog(*(c if withwt else c[:3]), bins=b) for c in cats]
    got = F.count_pairs(g[0], g[1] if len(g) > 1 else None, b, withwt=withwt)
    st = F.stats()
    for h in g:
        h.destroy()
    F.set_option("defaults", 0)
    ob = oracle.setup(prec=prec[0], periodic=periodic, arith=arith, **kw)
    pc = [oracle.preprocess(ob, c if withwt else c[:3]) for c in cats]
    want = oracle.count(ob, pc[0], pc[1] if len(pc) > 1 else None, withwt=withwt)
    ok = np.allclose(got, want, rtol=1e-12, atol=0) if withwt else np.array_equal(got, want)
    bad += not ok
    print(f"{tag:40s} {prec:6s} arith={arith} wt={int(withwt)} {'OK ' if ok else 'MISMATCH'} sum={got.sum():.6g} evals={st['pair_evals']} dense_rows={st['dense_rows']}", flush=True)


for prec in ("float", "double"):
    for arith in (0, 1):
        check("box smu fast (dense cells, k=3)", True, prec, [cat], dict(box=120.0, bintype=1, smax=40.0, ds=1.0, nmu=30), False, arith, [("k", 3)])
        check("box iso fast, shallow stacks", True, prec, [cat], dict(box=120.0, bintype=0, smax=30.0, ds=1.5), False, arith, [("qdepth", 12), ("qkeep", 1)])
    check("box smu generic", True, prec, [cat], dict(box=120.0, bintype=1, smax=30.0, ds=1.0, nmu=20), False, 1, [("force_generic", 1)])
    check("box spi weighted", True, prec, [cat], dict(box=120.0, bintype=2, smax=20.0, ds=1.0, pmin=0.0, pmax=30.0, dpi=2.0), True, 1)
    check("box smu weighted cross", True, prec, [cat, box_catalog(n // 2, 120.0, 8)], dict(box=120.0, bintype=1, smax=30.0, ds=1.0, nmu=20), True, 0)
    check("box smu global histogram", True, prec, [cat], dict(box=120.0, bintype=1, smax=30.0, ds=1.0, nmu=20), False, 1, [("global_hist", 1)])
    check("survey spi weighted DR", False, prec, [D, R], dict(bintype=2, smax=40.0, ds=2.0, pmin=0.0, pmax=80.0, dpi=1.0), True, 0)
    check("survey smu RR", False, prec, [R], dict(bintype=1, smax=120.0, ds=4.0, nmu=40), False, 1)
    check("survey iso DD", False, prec, [D], dict(bintype=0, smax=150.0, ds=5.0), False, 0)
F.lib().fcfc_gpu_finalize()
print("mismatches:", bad)
sys.exit(1 if bad else 0)
