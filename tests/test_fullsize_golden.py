"""BASELINE.json configs at their real size against the compiled reference.

tests/golden/fullsize_*.npz hold the raw count_pairs output of the unmodified reference (oracle/_ref) for
  c1   configs[0]: 10^6 uniform points, L = 1000, xi(s), 40 bins
  c2   configs[1]: 10^7 uniform points, L = 2000, xi(s,mu), 40 x 120 bins   (the bench workload)
  c4s  2x10^6 clustered points, xi(s,mu): the small-scale twin of configs[3]
  c4   10^7 clustered points, L = 2000, xi(s,mu): configs[3] at its single-GPU size (bench.py workload c4_box_smu_clustered_1e7)
  c3   configs[2]: survey, 2x10^6 data + 2x10^7 randoms, weighted DD + DR + RR, xi(s_perp,pi) 20 x 80 bins (double AVX-512 build)
generated once by tests/golden/make_golden_fullsize.py (tens of CPU minutes, hence fixtures).

Bars (north_star): double precision -- the reference's default build -- bit-exact in BOTH evaluation orders against the
reference's AVX-512 double build; single precision within the spread of the reference's own SINGLE_PREC builds
(k-d vs ball tree, scalar vs AVX-512: SURVEY.md section 7 measured that they disagree with each other in 34-36 of 40
bins, so no single float answer exists to be equal to)."""
import os

import numpy as np
import pytest

import bench
from conftest import GOLDEN

NAMES = ["c1", "c2", "c4s", "c4"]


def load(name):
    path = os.path.join(GOLDEN, f"fullsize_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


def float_refs(g):
    return {k[:-3]: g[k] for k in g.files if k.startswith("flt_") and k.endswith("_DD")}


# ------------------------------------------------------------------------------------------------- CPU: the fixtures
@pytest.mark.parametrize("name", NAMES)
def test_fixture_facts(name):
    """What the reference says about itself at full size: every double build agrees in every bin; the float builds
    do not, and their pairwise differences define the spread used below."""
    g = load(name)
    dbl = [g[k] for k in g.files if k.startswith("dbl_") and k.endswith("_DD")]
    for d in dbl[1:]:
        np.testing.assert_array_equal(d, dbl[0])
    fl = float_refs(g)
    assert len(fl) >= 2
    tot = int(dbl[0].sum())
    n, L = int(g["n"]), float(g["box"])
    if name not in ("c4s", "c4"):   # uniform catalogues: the analytic pair count, 5 sigma (+ the mu = 1 pairs dropped by (s,mu))
        expect = 0.5 * n * (n - 1) * (4.0 / 3.0 * np.pi * 200.0 ** 3) / L ** 3
        assert abs(tot - expect) < 5 * np.sqrt(expect) + 1e-5 * expect
    for v in fl.values():
        assert abs(int(v.sum()) - tot) < 1e-6 * tot


def reference_spread(g):
    """Per-bin range (max - min) over the reference's float builds."""
    fl = np.stack(list(float_refs(g).values()))
    return fl.max(axis=0) - fl.min(axis=0)


# ------------------------------------------------------------------------------------------------- GPU
def run_gpu(F, g, prec, arith):
    wl = bench.WORKLOADS[str(g["workload"])]
    n, L = int(g["n"]), float(g["box"])
    cat = bench.make_catalog(wl, n, L)
    kw = dict(bintype=wl["bintype"], smin=0.0, smax=200.0, ds=5.0)
    if wl["bintype"] == 1:
        kw["nmu"] = wl["nmu"]
    b = F.Bins(periodic=True, prec=prec, arith=arith, box=L, **kw)
    h = F.Catalog(*cat, bins=b)
    c = F.count_pairs(h, None, b)
    h.destroy()
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("arith", [0, 1])
def test_gpu_double_equals_reference_at_full_size(gpu, name, arith):
    g = load(name)
    got = run_gpu(gpu, g, "double", arith)
    np.testing.assert_array_equal(got, g["dbl_avx512_kd_DD"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("arith", [0, 1])
def test_gpu_float_within_reference_spread_at_full_size(gpu, name, arith):
    g = load(name)
    got = run_gpu(gpu, g, "float", arith)
    own = reference_spread(g)
    # the closest of the reference's float builds: scalar order vs its scalar build, FMA order vs its AVX-512 build
    ref = g["flt_scalar_kd_DD"] if arith == 0 else g["flt_avx512_kd_DD"]
    ours = np.abs(got - ref)
    print(f"{name} arith={arith}: reference's own float spread max {own.max()} sum {own.sum()} | engine vs "
          f"{'scalar' if arith == 0 else 'avx512'} k-d build: max {ours.max()} sum {ours.sum()} | total {int(got.sum())} vs {int(ref.sum())}")
    assert ours.max() <= max(4, 2 * own.max()) and ours.sum() <= max(16, 2 * own.sum())
    assert abs(int(got.sum()) - int(ref.sum())) <= max(4, 2 * int(np.abs(np.diff([int(v.sum()) for v in float_refs(g).values()])).max()))
    # and against double: float rounding moves a pair across an s or mu bin edge with probability ~1e-5
    assert np.abs(got - g["dbl_avx512_kd_DD"]).sum() <= 1e-4 * got.sum()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "golden", "fullsize_c3.npz")), reason="fullsize_c3.npz not generated")
def test_gpu_survey_weighted_equals_reference_at_full_size(gpu):
    """BASELINE configs[2] at its real size: 2x10^6 data + 2x10^7 randoms, weighted DD + DR + RR, xi(s_perp,pi) 20 x 80 bins,
    double precision, against the sums of the reference's default (double, AVX-512) build for the same catalogues:
    1e-12 relative per bin, identical empty bins."""
    import bench
    wl = bench.SURVEY_WORKLOADS["c3_svy_spi_wt_2e6_2e7"]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fullsize_c3.npz"))
    D, R = bench.make_survey(wl["nd"], 1), bench.make_survey(wl["nr"], 2)
    b = gpu.Bins(periodic=False, prec="double", bintype=2, smin=0.0, smax=wl["smax"], ds=wl["ds"], pmin=0.0, pmax=wl["pmax"], dpi=wl["dpi"])
    gd, gr = gpu.Catalog(*D, bins=b), gpu.Catalog(*R, bins=b)
    for name, a, c in (("DD", gd, None), ("DR", gd, gr), ("RR", gr, None)):
        got = gpu.count_pairs(a, c, b, withwt=True)
        ref = g[f"dbl_avx512_kd_{name}"]
        np.testing.assert_array_equal(got == 0, ref == 0)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0)
    gd.destroy(); gr.destroy()
