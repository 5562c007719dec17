"""Development parity sweep: GPU engine vs brute-force oracle vs compiled reference (run under gpurun)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcfc_b200 as F
from oracle import oracle, refdrv

rng = np.random.default_rng(20261017)
F.init(verbose=1)
nfail = 0

def survey_cat(n, rng):
    ra = np.deg2rad(rng.uniform(120, 180, n)); sd = rng.uniform(0, 0.5, n); cd = np.sqrt(1 - sd * sd)
    d = rng.uniform(1000, 1700, n)
    return (d * cd * np.cos(ra), d * cd * np.sin(ra), d * sd, rng.uniform(0.75, 1.25, n))

def check(name, periodic, prec, cats, pairs, withwt, arith=0, ref=False, **kw):
    global nfail
    b = F.Bins(periodic=periodic, prec=prec, arith=arith, **kw)
    ob = oracle.setup(prec=prec[0], periodic=periodic, arith=arith, **kw)
    g = [F.Catalog(*(c[:4] if withwt else c[:3]), bins=b) for c in cats]
    pc = [oracle.preprocess(ob, c if withwt else c[:3]) for c in cats]
    rr = None
    if ref:
        okw = dict(kw)
        rr = refdrv.run_reference([c if withwt else c[:3] for c in cats], periodic=periodic, prec='flt' if prec == 'float' else 'dbl',
                                  isa='scalar', pairs=[p for p in pairs], **okw)
    for k, p in enumerate(pairs):
        i, j = "DR".index(p[0]), "DR".index(p[1])
        t0 = time.time()
        c = F.count_pairs(g[i], None if i == j else g[j], b, withwt=withwt)
        st = F.stats()
        o = oracle.count(ob, pc[i], None if i == j else pc[j], withwt=withwt)
        if withwt:
            rel = np.max(np.abs(c - o) / np.maximum(np.abs(o), 1e-300)) if o.any() else 0
            ok = rel < 1e-12 and np.array_equal(c == 0, o == 0)
            msg = f"maxrel={rel:.2e}"
        else:
            ok = np.array_equal(c, o)
            msg = f"sum={c.sum()} diff={np.abs(c - o).sum()}"
        if rr is not None:
            rc = rr.pairs[k].cnt
            if withwt:
                rel2 = np.max(np.abs(c - rc) / np.maximum(np.abs(rc), 1e-300)) if rc.any() else 0
                ok = ok and rel2 < 1e-12
                msg += f" ref_maxrel={rel2:.2e}"
            else:
                ok = ok and np.array_equal(c, rc)
                msg += f" refdiff={np.abs(c - rc).sum()}"
        print(("PASS" if ok else "FAIL"), name, prec, p, msg, "ncell", st["ncell"], "items", st["nitem"], f"evals={st['pair_evals']:.3g} ms={st['ms_count']:.2f}", flush=True)
        if not ok: nfail += 1

N = 6000
x = rng.random((N, 3)) * 1000; y = rng.random((4000, 3)) * 1000
cat = (x[:, 0], x[:, 1], x[:, 2], rng.uniform(0.75, 1.25, N)); cat2 = (y[:, 0], y[:, 1], y[:, 2], rng.uniform(0.75, 1.25, 4000))
D = survey_cat(5000, rng); R = survey_cat(8000, rng)
for prec in ("double", "float"):
    for arith in (0, 1):
        tag = f"a{arith}"
        check("box-iso-" + tag, True, prec, [cat, cat2], ["DD", "DR"], False, arith, ref=(arith == 0), box=1000., bintype=0, smax=200., ds=5.)
        check("box-smu-" + tag, True, prec, [cat, cat2], ["DD", "DR"], False, arith, ref=(arith == 0), box=1000., bintype=1, smax=200., ds=5., nmu=120)
        check("box-spi-" + tag, True, prec, [cat, cat2], ["DD", "DR"], False, arith, ref=(arith == 0), box=1000., bintype=2, smax=100., ds=5., pmin=0., pmax=120., dpi=4.)
        check("box-iso-wt-" + tag, True, prec, [cat, cat2], ["DD", "DR"], True, arith, ref=(arith == 0), box=1000., bintype=0, smax=200., ds=5.)
        check("box-iso-smin-" + tag, True, prec, [cat], ["DD"], False, arith, ref=(arith == 0), box=1000., bintype=0, smin=10., smax=150., ds=2.5)
        check("box-smu-hyb-" + tag, True, prec, [cat], ["DD"], False, arith, ref=(arith == 0), box=1000., bintype=1, sbin_edges=np.logspace(0, np.log10(180), 21), nmu=50)
        check("box-spi-hyb-" + tag, True, prec, [cat], ["DD"], False, arith, ref=(arith == 0), box=1000., bintype=2, sbin_edges=np.logspace(0, np.log10(180), 21), pbin_edges=np.linspace(0, 99.5, 31))
        check("svy-iso-" + tag, False, prec, [D, R], ["DD", "DR", "RR"], False, arith, ref=(arith == 0), bintype=0, smax=200., ds=5.)
        check("svy-smu-wt-" + tag, False, prec, [D, R], ["DD", "DR"], True, arith, ref=(arith == 0), bintype=1, smax=200., ds=5., nmu=100)
        check("svy-smu-" + tag, False, prec, [D, R], ["DD", "DR"], False, arith, ref=(arith == 0), bintype=1, smax=200., ds=5., nmu=100)
        check("svy-spi-wt-" + tag, False, prec, [D, R], ["DD", "DR"], True, arith, ref=(arith == 0), bintype=2, smax=40., ds=2., pmin=0., pmax=80., dpi=1.)
        check("svy-spi-min-" + tag, False, prec, [D, R], ["DD", "DR"], False, arith, ref=(arith == 0), bintype=2, smin=4., smax=80., ds=2., pmin=10., pmax=150., dpi=2.5)
        check("svy-spi-hyb-" + tag, False, prec, [D], ["DD"], False, arith, ref=(arith == 0), bintype=2, sbin_edges=np.logspace(-1, 2, 16), pbin_edges=np.linspace(0, 120, 25))
print("FAILURES:", nfail)

# quick timing: C1-like
if len(sys.argv) > 1:
    for (N, L, bt, nmu) in ((1000000, 1000., 0, 1), (1000000, 1000., 1, 120), (10000000, 2000., 1, 120)):
        x = rng.random((N, 3)) * L
        for prec in ("float", "double"):
            for arith in (1, 0):
                b = F.Bins(periodic=True, prec=prec, arith=arith, box=L, bintype=bt, smax=200., ds=5., nmu=nmu)
                t0 = time.time(); g = F.Catalog(x[:, 0], x[:, 1], x[:, 2], bins=b); t1 = time.time()
                c = F.count_pairs(g, None, b); st = F.stats(); t2 = time.time()
                c = F.count_pairs(g, None, b); st2 = F.stats()
                print(f"N={N} bt={bt} {prec} arith={arith}: upload {t1 - t0:.3f}s count {t2 - t1:.3f}s kernel {st['ms_count']:.1f}/{st2['ms_count']:.1f} ms sort {st['ms_sort']:.1f} ms evals {st['pair_evals']:.4g} pairs {st['pairs_in']:.5g} -> {st2['pair_evals'] / st2['ms_count'] * 1e-9:.3f} Tevals/s ncell {st['ncell']} items {st['nitem']}", flush=True)
                g.destroy()
sys.exit(1 if nfail else 0)
