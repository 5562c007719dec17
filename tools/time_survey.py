"""Timing of survey-mode weighted counts (configs[2]-like, scaled): python tools/time_survey.py [nd] [nr] [prec]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fcfc_b200 as F
from cases import survey_catalog
nd = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000
nr = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2000000
prec = sys.argv[3] if len(sys.argv) > 3 else "double"
withref = len(sys.argv) > 4
F.init()
D, R = survey_catalog(nd, 1), survey_catalog(nr, 2)
only = [int(v) for v in os.environ.get("FCFC_TS_BINTYPES", "2,1,0").split(",")]      # restrict the binning schemes (large catalogues)
for bt, kw in ((2, dict(smax=40., ds=2., pmin=0., pmax=80., dpi=1.)), (1, dict(smax=200., ds=5., nmu=120)), (0, dict(smax=200., ds=5.))):
    if bt not in only:
        continue
    for wt in ((True,) if os.environ.get("FCFC_TS_WEIGHTED_ONLY") else (True, False)):
        b = F.Bins(periodic=False, prec=prec, bintype=bt, arith=0, **kw)
        gd = F.Catalog(*(D if wt else D[:3]), bins=b); gr = F.Catalog(*(R if wt else R[:3]), bins=b)
        out = []
        for name, a, c in (("DD", gd, None), ("DR", gd, gr), ("RR", gr, None)):
            F.count_pairs(a, c, b, withwt=wt)
            cnt = F.count_pairs(a, c, b, withwt=wt); st = F.stats()
            out.append(f"{name} {st['ms_count']:.1f} ms ({st['pair_evals']:.3g} evals, sum {cnt.sum():.6g}, grid {st['ncell']})")
        print(f"bintype={bt} weighted={wt} {prec}: " + "; ".join(out), flush=True)
        gd.destroy(); gr.destroy()
if withref:
    from oracle import refdrv
    p = "flt" if prec == "float" else "dbl"
    isa = refdrv.best_simd_flavour(p).split("_")[1]
    r = refdrv.run_reference([D, R], periodic=False, prec=p, isa=isa, pairs=["DD", "DR", "RR"], bintype=2, smax=40., ds=2., pmin=0., pmax=80., dpi=1.)
    print("reference", p, isa, [(q.label, round(q.t_count, 3)) for q in r.pairs], "threads", r.nthread)
