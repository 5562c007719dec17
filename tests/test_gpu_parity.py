"""GPU (-m gpu): the CUDA engine, called through the C ABI, against
  (1) the golden vectors of the compiled reference (tests/golden),
  (2) the brute-force CPU restatement (oracle/) on seeded inputs,
  (3) the compiled reference itself run live on the host CPU (oracle/_ref travels to the GPU box),
  (4) size-independent properties at BASELINE.json's full sizes.
Bars: unweighted counts bit-exact; weighted FP64 sums within 1e-12 relative (north_star)."""
import numpy as np
import pytest

from cases import CASES, box_catalog, clustered_box_catalog, lattice_box_catalog, make_catalog, survey_catalog
from conftest import have_ref, load_golden
from oracle import oracle, refdrv

pytestmark = pytest.mark.gpu

WT_RTOL = 1e-12


def gpu_counts(F, case_kw, periodic, prec, cats, pairs, withwt, arith=0):
    b = F.Bins(periodic=periodic, prec=prec, arith=arith, **case_kw)
    g = [F.Catalog(*c, bins=b) for c in cats]
    out = {}
    for p in pairs:
        i, j = "DR".index(p[0]), "DR".index(p[1])
        out[p] = F.count_pairs(g[i], None if i == j else g[j], b, withwt=withwt)
    for c in g:
        c.destroy()
    return out


def assert_counts(got, want, withwt):
    if withwt:
        np.testing.assert_array_equal(got == 0, want == 0)
        np.testing.assert_allclose(got, want, rtol=WT_RTOL, atol=0)
    else:
        np.testing.assert_array_equal(got, want)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("prec", ["dbl", "flt"])
def test_golden_reference_vectors(gpu, name, prec):
    case = CASES[name]
    g = load_golden(name)
    cats = [make_catalog(s, case["withwt"]) for s in case["cats"]]
    got = gpu_counts(gpu, case["kw"], case["periodic"], "float" if prec == "flt" else "double", cats,
                     case["pairs"], case["withwt"])
    for p in case["pairs"]:
        assert_counts(got[p], g[f"{prec}_scalar_{p}"], case["withwt"])
        if prec == "dbl" and not case["withwt"]:
            np.testing.assert_array_equal(got[p], g[f"dbl_avx512_{p}"])     # default (SIMD) build of the reference


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("prec", ["double", "float"])
@pytest.mark.parametrize("arith", [0, 1])
def test_against_oracle(gpu, name, prec, arith):
    """Both arithmetic orders (scalar parity mode, FMA mode) against the brute-force restatement."""
    case = CASES[name]
    cats = [make_catalog(s, case["withwt"]) for s in case["cats"]]
    got = gpu_counts(gpu, case["kw"], case["periodic"], prec, cats, case["pairs"], case["withwt"], arith)
    ob = oracle.setup(prec=prec[0], periodic=case["periodic"], arith=arith, **case["kw"])
    pc = [oracle.preprocess(ob, c) for c in cats]
    for p in case["pairs"]:
        i, j = "DR".index(p[0]), "DR".index(p[1])
        want = oracle.count(ob, pc[i], None if i == j else pc[j], withwt=case["withwt"])
        assert_counts(got[p], want, case["withwt"])


@pytest.mark.parametrize("prec", ["double", "float"])
@pytest.mark.parametrize("kw", [dict(bintype=1, smax=40.0, ds=1.0, nmu=120), dict(bintype=0, smax=37.5, ds=2.5),
                                dict(bintype=2, smax=30.0, ds=1.0, pmin=0.0, pmax=40.0, dpi=1.0)])
def test_medium_box_vs_oracle(gpu, prec, kw):
    """25k points in a small box: many cells, several tiles per cell, every periodic image exercised."""
    cat = box_catalog(25000, 300.0, 31, weights=False)
    got = gpu_counts(gpu, dict(box=300.0, **kw), True, prec, [cat], ["DD"], False)["DD"]
    ob = oracle.setup(prec=prec[0], periodic=True, box=300.0, **kw)
    np.testing.assert_array_equal(got, oracle.count(ob, oracle.preprocess(ob, cat)))


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("kw", [dict(bintype=1, smax=42.0, ds=1.0, nmu=50), dict(bintype=0, smax=42.0, ds=1.5)])
def test_dense_cells_vs_oracle(gpu, kw, arith, tuned):
    """The dense-cell path (secondary cells entirely in range are binned in place by the pair loop, count_kernel.cuh:
    do_chunk_dense) needs full tiles: 25k points in a box of 7^3 cells of a third of the maximum separation (forced
    with the engine option k) give ~73 points per cell and seven dense offsets.  Counts must equal the oracle's and those of
    the same engine with the path switched off; auto and cross counts."""
    a, b = box_catalog(25000, 100.0, 41, weights=False), box_catalog(20000, 100.0, 42, weights=False)
    tuned("k", 3)
    bins = gpu.Bins(periodic=True, prec="float", arith=arith, box=100.0, **kw)
    ga, gb = gpu.Catalog(*a, bins=bins), gpu.Catalog(*b, bins=bins)
    on = {}
    for name, c2 in (("DD", None), ("DR", gb)):
        on[name] = gpu.count_pairs(ga, c2, bins)
        assert gpu.stats()["dense_rows"] > 0, "the dense path was not used"
    tuned("no_dense", 1)
    for name, c2 in (("DD", None), ("DR", gb)):
        off = gpu.count_pairs(ga, c2, bins)
        assert gpu.stats()["dense_rows"] == 0
        np.testing.assert_array_equal(on[name], off)
    ga.destroy(); gb.destroy()
    ob = oracle.setup(prec="f", periodic=True, arith=arith, box=100.0, **kw)
    pa, pb = oracle.preprocess(ob, a), oracle.preprocess(ob, b)
    np.testing.assert_array_equal(on["DD"], oracle.count(ob, pa))
    np.testing.assert_array_equal(on["DR"], oracle.count(ob, pa, pb))


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("kw", [dict(bintype=1, smax=30.0, ds=1.0, nmu=40), dict(bintype=0, smax=33.0, ds=1.5)])
@pytest.mark.parametrize("origin", [0.0, -37.5])
def test_classified_staging_vs_plain(gpu, kw, arith, origin, tuned):
    """Single-precision box counts take count_kernel_cl (staged points classified against the tile's bounding box, two
    compacted rings, tiles of 96 points): the statistics must say so -- points were dropped without a distance evaluation --
    and the counts must equal those of the plain kernel (option no_classify) and of the oracle.  36k points in a 120 box
    (reach a quarter of it): every periodic image, several tiles per cell in the clustered part, auto and cross."""
    a = box_catalog(24000, 120.0, 61, weights=False)
    c = clustered_box_catalog(12000, 120.0, 62)[:3]
    a = tuple(np.concatenate([u, v]) + origin for u, v in zip(a, c))
    b = tuple(u + origin for u in box_catalog(9000, 120.0, 63, weights=False))
    bins = gpu.Bins(periodic=True, prec="float", arith=arith, box=120.0, **kw)
    ga, gb = gpu.Catalog(*a, bins=bins), gpu.Catalog(*b, bins=bins)
    on = {}
    for name, c2 in (("DD", None), ("DR", gb)):
        on[name] = gpu.count_pairs(ga, c2, bins)
        st = gpu.stats()
        assert st["classified"] == 1 and st["pair_evals_computed"] < st["pair_evals"], "classified staging was not used"
    tuned("no_classify", 1)
    for name, c2 in (("DD", None), ("DR", gb)):
        off = gpu.count_pairs(ga, c2, bins)
        st = gpu.stats()
        assert st["classified"] == 0 and st["pair_evals_computed"] == st["pair_evals"]
        np.testing.assert_array_equal(on[name], off)
    ga.destroy(); gb.destroy()
    if origin == 0.0:
        ob = oracle.setup(prec="f", periodic=True, arith=arith, box=120.0, **kw)
        pa, pb = oracle.preprocess(ob, a), oracle.preprocess(ob, b)
        np.testing.assert_array_equal(on["DD"], oracle.count(ob, pa))
        np.testing.assert_array_equal(on["DR"], oracle.count(ob, pa, pb))


@pytest.mark.parametrize("arith", [0, 1])
def test_survey_isotropic_float_classified(gpu, arith, tuned):
    """Survey isotropic counts in single precision take count_kernel_cl as well (no image shifts: the rings are only flushed at
    the end of a work item); against the plain kernel and the oracle, auto and cross."""
    D, R = survey_catalog(20000, 91)[:3], survey_catalog(30000, 92)[:3]
    kw = dict(bintype=0, smax=150.0, ds=5.0)
    bins = gpu.Bins(periodic=False, prec="float", arith=arith, **kw)
    gd, gr = gpu.Catalog(*D, bins=bins), gpu.Catalog(*R, bins=bins)
    on = {}
    for name, a, c in (("DD", gd, None), ("DR", gd, gr)):
        on[name] = gpu.count_pairs(a, c, bins)
        st = gpu.stats()
        assert st["classified"] == 1 and st["pair_evals_computed"] < st["pair_evals"]
    tuned("no_classify", 1)
    for name, a, c in (("DD", gd, None), ("DR", gd, gr)):
        np.testing.assert_array_equal(on[name], gpu.count_pairs(a, c, bins))
        assert gpu.stats()["classified"] == 0
    gd.destroy(); gr.destroy()
    ob = oracle.setup(prec="f", periodic=False, arith=arith, **kw)
    pd, pr = oracle.preprocess(ob, D), oracle.preprocess(ob, R)
    np.testing.assert_array_equal(on["DD"], oracle.count(ob, pd))
    np.testing.assert_array_equal(on["DR"], oracle.count(ob, pd, pr))


@pytest.mark.parametrize("prec", ["float", "double"])
@pytest.mark.parametrize("kw", [dict(bintype=1, smax=49.0, ds=3.5, nmu=15), dict(bintype=0, smax=49.5, ds=1.5)])
def test_reach_of_almost_half_the_box(gpu, prec, kw):
    """Maximum separation just below half the box (the reference's own limit): every cell is a neighbour of every other
    through some periodic image, a stencil row wraps around and meets itself, image shifts change from row to row -- the
    bookkeeping the classified staging has to flush its rings for.  12k points, auto and cross, against the oracle."""
    L = 100.0
    a = box_catalog(9000, L, 81, weights=False)
    c = clustered_box_catalog(3000, L, 82)[:3]
    a = tuple(np.concatenate([u, v]) for u, v in zip(a, c))
    b = box_catalog(5000, L, 83, weights=False)
    got = gpu_counts(gpu, dict(box=L, **kw), True, prec, [a, b], ["DD", "DR"], False, arith=1)
    ob = oracle.setup(prec=prec[0], periodic=True, arith=1, box=L, **kw)
    pa, pb = oracle.preprocess(ob, a), oracle.preprocess(ob, b)
    np.testing.assert_array_equal(got["DD"], oracle.count(ob, pa))
    np.testing.assert_array_equal(got["DR"], oracle.count(ob, pa, pb))


@pytest.mark.parametrize("kw,withwt", [(dict(bintype=2, smax=24.0, ds=2.0, pmin=0.0, pmax=60.0, dpi=2.0), True),
                                       (dict(bintype=2, smax=40.0, ds=2.0, pmin=0.0, pmax=30.0, dpi=1.0), False),
                                       (dict(bintype=1, smax=60.0, ds=3.0, nmu=30), True)])
def test_survey_classification_vs_plain_double(gpu, kw, withwt, tuned):
    """Double-precision survey counts take the pre-filter kernel with the per-point classification (sphere and, for
    (s_perp, pi), the two cylinder bounds) and the dealt exact pass: results must equal the plain FP64 kernel's
    (option no_prefilter) -- unweighted bit for bit, weighted to 1e-12 with the same empty bins."""
    D, R = survey_catalog(30000, 71), survey_catalog(90000, 72)
    cats = [D, R] if withwt else [D[:3], R[:3]]
    bins = gpu.Bins(periodic=False, prec="double", **kw)
    g = [gpu.Catalog(*c, bins=bins) for c in cats]
    got = {}
    for name, i, j in (("DD", 0, 0), ("DR", 0, 1), ("RR", 1, 1)):
        got[name] = gpu.count_pairs(g[i], None if i == j else g[j], bins, withwt=withwt)
        st = gpu.stats()
        assert st["prefilter"] == 1 and st["classified"] == 1 and st["pair_evals_computed"] < st["pair_evals"]
    tuned("no_prefilter", 1)
    for name, i, j in (("DD", 0, 0), ("DR", 0, 1), ("RR", 1, 1)):
        plain = gpu.count_pairs(g[i], None if i == j else g[j], bins, withwt=withwt)
        assert gpu.stats()["prefilter"] == 0
        assert_counts(got[name], plain, withwt)
    for c in g:
        c.destroy()


@pytest.mark.parametrize("prec", ["double", "float"])
def test_clustered_cuboid_vs_oracle(gpu, prec):
    cat = clustered_box_catalog(20000, 400.0, 32)
    x, y, z, w = cat
    cat = (x, y * 0.75, z * 0.5, w)                     # cuboid box 400 x 300 x 200
    kw = dict(box=[400.0, 300.0, 200.0], bintype=1, smax=30.0, ds=1.5, nmu=40)
    for withwt in (False, True):
        c = cat if withwt else cat[:3]
        got = gpu_counts(gpu, kw, True, prec, [c], ["DD"], withwt)["DD"]
        ob = oracle.setup(prec=prec[0], periodic=True, **kw)
        assert_counts(got, oracle.count(ob, oracle.preprocess(ob, c), withwt=withwt), withwt)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not have_ref("dbl_scalar", "box"), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("kw", [dict(bintype=0, smax=200.0, ds=5.0), dict(bintype=1, smax=200.0, ds=5.0, nmu=120)])
def test_live_reference_box_double(gpu, kw):
    """2x10^5 points against the unmodified reference (its fastest build on this host), bit-exact."""
    cat = box_catalog(200000, 1000.0, 41, weights=False)
    flavour = refdrv.best_simd_flavour("dbl").split("_")[1]
    r = refdrv.run_reference([cat], periodic=True, prec="dbl", isa=flavour, pairs=["DD"], box=1000.0, **kw)
    got = gpu_counts(gpu, dict(box=1000.0, **kw), True, "double", [cat], ["DD"], False)["DD"]
    np.testing.assert_array_equal(got, r.pairs[0].cnt)


@pytest.mark.skipif(not have_ref("flt_scalar", "svy"), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("prec", ["dbl", "flt"])
def test_live_reference_survey_weighted(gpu, prec):
    """Survey mode never crosses a periodic boundary, so SINGLE_PREC is exact against the scalar reference too."""
    D, R = survey_catalog(40000, 42), survey_catalog(120000, 43)
    kw = dict(bintype=2, smax=40.0, ds=2.0, pmin=0.0, pmax=80.0, dpi=1.0)
    r = refdrv.run_reference([D, R], periodic=False, prec=prec, isa="scalar", pairs=["DD", "DR", "RR"], **kw)
    got = gpu_counts(gpu, kw, False, "float" if prec == "flt" else "double", [D, R], ["DD", "DR", "RR"], True)
    for k, p in enumerate(["DD", "DR", "RR"]):
        assert_counts(got[p], r.pairs[k].cnt, True)
    r = refdrv.run_reference([D[:3], R[:3]], periodic=False, prec=prec, isa="scalar", pairs=["DD", "DR"], **kw)
    got = gpu_counts(gpu, kw, False, "float" if prec == "flt" else "double", [D[:3], R[:3]], ["DD", "DR"], False)
    for k, p in enumerate(["DD", "DR"]):
        np.testing.assert_array_equal(got[p], r.pairs[k].cnt)


@pytest.mark.skipif(not have_ref("flt_scalar", "box"), reason="oracle/_ref not shipped")
def test_live_reference_box_float_spread(gpu):
    """Gate G2: against the SINGLE_PREC reference on a generic periodic catalogue the per-bin difference must
    stay within the reference's own k-d-tree vs ball-tree spread (which is reported, not hidden)."""
    cat = box_catalog(200000, 1000.0, 44, weights=False)
    kw = dict(bintype=0, smax=200.0, ds=5.0)
    kd = refdrv.run_reference([cat], periodic=True, prec="flt", isa="scalar", pairs=["DD"], box=1000.0, data_struct=0, **kw).pairs[0].cnt
    ball = refdrv.run_reference([cat], periodic=True, prec="flt", isa="scalar", pairs=["DD"], box=1000.0, data_struct=1, **kw).pairs[0].cnt
    got = gpu_counts(gpu, dict(box=1000.0, **kw), True, "float", [cat], ["DD"], False)["DD"]
    own = np.abs(kd - ball)
    ours = np.abs(got - kd)
    print("reference kd-vs-ball: max", own.max(), "sum", own.sum(), "| engine-vs-kd: max", ours.max(), "sum", ours.sum())
    assert ours.max() <= max(4, 2 * own.max()) and ours.sum() <= max(16, 2 * own.sum())
    assert abs(int(got.sum()) - int(kd.sum())) <= max(4, 2 * abs(int(kd.sum()) - int(ball.sum())))


# ---------------------------------------------------------------------------------------------------
def test_edge_cases(gpu):
    F = gpu
    kw = dict(box=100.0, bintype=1, smax=20.0, ds=1.0, nmu=10)
    for prec in ("double", "float"):
        b = F.Bins(periodic=True, prec=prec, **kw)
        ob = oracle.setup(prec=prec[0], periodic=True, **kw)
        empty = F.Catalog([], [], [], bins=b)
        one = F.Catalog([5.0], [5.0], [5.0], bins=b)
        assert F.count_pairs(empty, None, b).sum() == 0
        assert F.count_pairs(one, None, b).sum() == 0
        assert F.count_pairs(one, empty, b).sum() == 0
        # coincident points: every distinct pair lands in bin 0 (d^2 < EPS -> mu index 0)
        n = 70
        same = F.Catalog([50.0] * n, [50.0] * n, [50.0] * n, bins=b)
        c = F.count_pairs(same, None, b)
        assert c[0] == n * (n - 1) // 2 and c.sum() == c[0]
        # ragged sizes around the tile (128) and warp (32) boundaries, points on the box faces
        for n in (2, 31, 33, 127, 128, 129, 257, 1000):
            rng = np.random.default_rng(n)
            x = rng.random((n, 3)) * 100.0
            x[0] = [0.0, 0.0, 0.0]
            x[1] = [100.0, 100.0, 100.0] if prec == "float" else [99.999999, 0.0, 50.0]
            cat = (x[:, 0], x[:, 1], x[:, 2])
            g = F.Catalog(*cat, bins=b)
            np.testing.assert_array_equal(F.count_pairs(g, None, b), oracle.count(ob, oracle.preprocess(ob, cat)))
            np.testing.assert_array_equal(F.count_pairs(g, one, b),
                                          oracle.count(ob, oracle.preprocess(ob, cat), oracle.preprocess(ob, ([5.0], [5.0], [5.0]))))


@pytest.mark.parametrize("prec", ["double", "float"])
@pytest.mark.parametrize("origin", [-150.0, -300.0, 1234.5, -1e-7])
def test_box_with_any_origin(gpu, prec, origin):
    """The reference never range-checks positions: only differences and +-L shifts enter its metric
    (metric_kdtree.c:55-67, metric_common.c:998), so a box stored as [-L/2, L/2) -- or starting anywhere -- counts
    like one stored as [0, L).  The cell grid starts at the catalogues' common minimum in that case."""
    L = 300.0
    x, y, z = box_catalog(20000, L, 91, weights=False)
    a = (x + origin, y + origin, z + 0.5 * origin)
    b = tuple(np.ascontiguousarray(c[:7000][::-1]) for c in a)
    kw = dict(box=L, bintype=1, smax=30.0, ds=1.5, nmu=25)
    got = gpu_counts(gpu, kw, True, prec, [a, b], ["DD", "DR"], False)
    ob = oracle.setup(prec=prec[0], periodic=True, **kw)
    pa, pb = oracle.preprocess(ob, a), oracle.preprocess(ob, b)
    np.testing.assert_array_equal(got["DD"], oracle.count(ob, pa))
    np.testing.assert_array_equal(got["DR"], oracle.count(ob, pa, pb))
    if prec == "double" and origin == -150.0:
        # points spread over more than one period are refused loudly
        F = gpu
        bins = F.Bins(periodic=True, prec=prec, **kw)
        bad = F.Catalog(np.append(a[0], origin + 1.5 * L), np.append(a[1], 0.0), np.append(a[2], 0.0), bins=bins)
        with pytest.raises(F.FcfcGpuError, match="outside one period"):
            F.count_pairs(bad, None, bins)


@pytest.mark.parametrize("prec", ["double", "float"])
def test_tiny_reach_in_huge_volume(gpu, prec):
    """s_max = 2 in a 20000 box (and a sparse survey volume): one reach per cell would need > 2048 cells per axis,
    the grid falls back to coarser cells.  Pairs are planted so that the histogram is not empty."""
    rng = np.random.default_rng(77)
    L = 20000.0
    x = rng.random((6000, 3)) * L
    x[1000:2000] = (x[:1000] + rng.normal(0, 0.6, (1000, 3))) % L      # close companions, some across the faces
    x[:50, 0] = rng.random(50) * 0.5; x[1000:1050, 0] = L - rng.random(50) * 0.5
    x[1000:1050, 1:] = x[:50, 1:]
    x = np.round(x, 5); x[x >= L] = 0.0
    cat = (x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy())
    kw = dict(box=L, bintype=1, smax=2.0, ds=0.25, nmu=20)
    got = gpu_counts(gpu, kw, True, prec, [cat], ["DD"], False)["DD"]
    ob = oracle.setup(prec=prec[0], periodic=True, **kw)
    want = oracle.count(ob, oracle.preprocess(ob, cat))
    assert want.sum() > 500
    np.testing.assert_array_equal(got, want)
    if prec == "float":
        return          # (s1 + s2) - t at |x| ~ 5e4 has no significant digits left in float for s < 2
    # survey geometry: same idea without periodic images
    sx, sy, sz, sw = survey_catalog(4000, 78)
    s = np.stack([sx, sy, sz], 1) * 40.0                               # 40000-68000 Mpc/h: enormous, sparse volume
    s[2000:] = s[:2000] + rng.normal(0, 0.5, (2000, 3))
    scat = (s[:, 0].copy(), s[:, 1].copy(), s[:, 2].copy(), sw)
    kw = dict(bintype=2, smax=2.0, ds=0.25, pmin=0.0, pmax=3.0, dpi=0.5)
    got = gpu_counts(gpu, kw, False, prec, [scat], ["DD"], True)["DD"]
    ob = oracle.setup(prec=prec[0], periodic=False, **kw)
    want = oracle.count(ob, oracle.preprocess(ob, scat), withwt=True)
    assert (want > 0).sum() > 20
    assert_counts(got, want, True)


@pytest.mark.parametrize("prec", ["double", "float"])
@pytest.mark.parametrize("arith", [0, 1])
def test_medium_survey_smu_vs_oracle(gpu, prec, arith):
    """Survey (s,mu) with computed bins (pi / s from reciprocal square roots, flagged pairs re-binned exactly):
    10^8 - 10^9 candidate pairs against the brute-force oracle, unweighted exact, weighted to 1e-12."""
    D, R = survey_catalog(16000, 51), survey_catalog(24000, 52)
    kw = dict(bintype=1, smax=120.0, ds=4.0, nmu=60)
    ob = oracle.setup(prec=prec[0], periodic=False, arith=arith, **kw)
    pD, pR = oracle.preprocess(ob, D), oracle.preprocess(ob, R)
    got = gpu_counts(gpu, kw, False, prec, [D[:3], R[:3]], ["DD", "DR"], False, arith)
    np.testing.assert_array_equal(got["DD"], oracle.count(ob, oracle.preprocess(ob, D[:3])))
    np.testing.assert_array_equal(got["DR"], oracle.count(ob, oracle.preprocess(ob, D[:3]), oracle.preprocess(ob, R[:3])))
    gotw = gpu_counts(gpu, kw, False, prec, [D, R], ["DR"], True, arith)["DR"]
    assert_counts(gotw, oracle.count(ob, pD, pR, withwt=True), True)


@pytest.mark.parametrize("depth,keep", [(8, 0), (8, 4), (12, 1), (16, 8), (24, 3), (64, 0)])
def test_stack_depth_and_keep(gpu, depth, keep, tuned):
    """The per-lane stacks at every depth the shared-memory plan can choose (shallow ones are what the weighted
    double-precision variants get), with drains that leave nothing / half of the stack: same counts."""
    tuned("qdepth", depth)
    tuned("qkeep", keep)
    cat = box_catalog(6000, 250.0, 33)
    for prec in ("float", "double"):
        for kw, withwt in ((dict(bintype=1, smax=40.0, ds=1.0, nmu=30), False), (dict(bintype=0, smax=40.0, ds=2.0), False),
                           (dict(bintype=2, smax=30.0, ds=1.0, pmin=0.0, pmax=40.0, dpi=2.0), True)):
            c = cat if withwt else cat[:3]
            got = gpu_counts(gpu, dict(box=250.0, **kw), True, prec, [c], ["DD"], withwt, arith=1)["DD"]
            ob = oracle.setup(prec=prec[0], periodic=True, arith=1, box=250.0, **kw)
            assert_counts(got, oracle.count(ob, oracle.preprocess(ob, c), withwt=withwt), withwt)
    D, R = survey_catalog(3000, 34), survey_catalog(4000, 35)
    kw = dict(bintype=1, smax=150.0, ds=5.0, nmu=20)
    got = gpu_counts(gpu, kw, False, "double", [D, R], ["DR"], True)["DR"]
    ob = oracle.setup(prec="d", periodic=False, **kw)
    assert_counts(got, oracle.count(ob, oracle.preprocess(ob, D), oracle.preprocess(ob, R), withwt=True), True)


def test_huge_histogram_uses_global_path(gpu):
    """ns = 600 x nmu = 255 = 153000 bins do not fit the shared-memory histogram: global-atomic variant."""
    kw = dict(box=500.0, bintype=1, smax=60.0, ds=0.1, nmu=255)
    cat = box_catalog(6000, 500.0, 51, weights=False)
    for prec in ("double", "float"):
        got = gpu_counts(gpu, kw, True, prec, [cat], ["DD"], False)["DD"]
        ob = oracle.setup(prec=prec[0], periodic=True, **kw)
        np.testing.assert_array_equal(got, oracle.count(ob, oracle.preprocess(ob, cat)))


def test_with_mu_one(gpu):
    kw = dict(box=200.0, bintype=1, smax=40.0, ds=2.0, nmu=20, with_mu_one=True)
    x, y, z, _ = lattice_box_catalog(3000, 200.0, 52)
    x[1000:2000], y[1000:2000] = x[:1000], y[:1000]     # 1000 pairs exactly along z (mu = 1)
    for prec in ("double", "float"):
        got = gpu_counts(gpu, kw, True, prec, [(x, y, z)], ["DD"], False)["DD"]
        ob = oracle.setup(prec=prec[0], periodic=True, **kw)
        np.testing.assert_array_equal(got, oracle.count(ob, oracle.preprocess(ob, (x, y, z))))
        kw0 = dict(kw, with_mu_one=False)
        got0 = gpu_counts(gpu, kw0, True, prec, [(x, y, z)], ["DD"], False)["DD"]
        assert got0.sum() < got.sum()


def test_errors(gpu):
    F = gpu
    b = F.Bins(periodic=True, prec="double", box=100.0, bintype=0, smax=20.0, ds=1.0)
    bad = F.Catalog([1.0, 120.0], [1.0, 2.0], [1.0, 2.0], bins=b)
    with pytest.raises(F.FcfcGpuError, match="outside one period of the box") as e:
        F.count_pairs(bad, None, b)
    assert e.value.code == -21
    bf = F.Bins(periodic=True, prec="float", box=100.0, bintype=0, smax=20.0, ds=1.0)
    with pytest.raises(F.FcfcGpuError, match="precision mismatch"):
        F.count_pairs(bad, None, bf)
    with pytest.raises(F.FcfcGpuError, match="non-finite"):
        F.Catalog([1.0, float("nan")], [1.0, 2.0], [1.0, 2.0], bins=b)
    with pytest.raises(F.FcfcGpuError, match="half the box"):
        F.count_pairs(F.Catalog([1.0], [1.0], [1.0], bins=b), None,
                      F.Bins(periodic=True, prec="double", box=30.0, bintype=0, smax=20.0, ds=1.0))


# ---------------------------------------------------------------------------------------------------
# Properties at BASELINE.json sizes (no brute force possible)
@pytest.fixture(scope="module")
def million():
    return box_catalog(1000000, 1000.0, 61, weights=False)


def test_full_size_properties_c1(gpu, million):
    """configs[0]: 10^6 points, L = 1000, s in [0, 200) in 40 bins."""
    F = gpu
    for prec in ("float", "double"):
        biso = F.Bins(periodic=True, prec=prec, box=1000.0, bintype=0, smax=200.0, ds=5.0)
        bsmu = F.Bins(periodic=True, prec=prec, box=1000.0, bintype=1, smax=200.0, ds=5.0, nmu=120, with_mu_one=True)
        g = F.Catalog(*million, bins=biso)
        iso = F.count_pairs(g, None, biso)
        st = F.stats()
        # idempotence
        np.testing.assert_array_equal(iso, F.count_pairs(g, None, biso))
        # shards add up (the multi-GPU decomposition), any number of parts
        for nparts in (2, 3, 8):
            tot = sum(F.count_pairs(g, None, biso, part=p, nparts=nparts) for p in range(nparts))
            np.testing.assert_array_equal(tot, iso)
        # (s, mu) with mu = 1 kept, summed over mu, is the isotropic count
        smu = F.count_pairs(g, None, bsmu)
        np.testing.assert_array_equal(smu.reshape(120, 40).sum(axis=0), iso)
        # cross count of a catalogue with a copy of itself = ordered pairs + the N self pairs in bin 0
        g2 = F.Catalog(*million, bins=biso)
        cross = F.count_pairs(g, g2, biso)
        want = 2 * iso
        want[0] += len(million[0])
        np.testing.assert_array_equal(cross, want)
        # expected number of pairs for a uniform field (4/3 pi r^3 n^2 / 2), 5 sigma
        n = len(million[0])
        expect = 0.5 * n * (n - 1) * (4.0 / 3.0 * np.pi * 200.0 ** 3) / 1000.0 ** 3
        assert abs(iso.sum() - expect) < 5 * np.sqrt(expect) + 1e-4 * expect
        assert st["pair_evals"] >= iso.sum()
        g.destroy(); g2.destroy()


def test_full_size_lattice_float_equals_double(gpu):
    """10^6 lattice points, power-of-two bins: every float operation is exact, so float == double exactly,
    in both arithmetic orders (gate G2's dyadic-grid check at full size)."""
    F = gpu
    cat = lattice_box_catalog(1000000, 1024.0, 62)[:3]
    kw = dict(box=1024.0, bintype=2, smax=128.0, ds=8.0, pmin=0.0, pmax=128.0, dpi=8.0)
    res = []
    for prec in ("double", "float"):
        for arith in (0, 1):
            res.append(gpu_counts(F, kw, True, prec, [cat], ["DD"], False, arith)["DD"])
    for r in res[1:]:
        np.testing.assert_array_equal(r, res[0])
