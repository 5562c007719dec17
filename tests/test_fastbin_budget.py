"""CPU: the error budget behind the computed ("fast") bins of the counting kernels (count_kernel.cuh: fast_bins).

The kernels bin a box / isotropic pair from ONE approximate reciprocal square root and flag every pair whose scaled
s or nmu*mu lands within a band of a bin edge; only flagged pairs are re-binned with the reference's exact IEEE
sequence.  Bit-exactness therefore rests on one claim: *an unflagged pair's fast bins equal the exact bins*.  This test
emulates the device arithmetic in numpy (float32 operations are IEEE round-to-nearest like the device's; rsqrt.approx
is modelled as the true value times 1 +- 2^-22, its documented bound, with random and extreme signs; the fused
multiply-add rounded toward zero is exact integer arithmetic here) on pairs drawn at random and adversarially close to
bin edges, with the scales 2^ks, 2^km taken from the library itself (fcfc_gpu_fastbin_scales)."""
import ctypes

import numpy as np
import pytest

import fcfc_b200 as F

f32 = np.float32


def scales(ns, nmu, periodic=1):
    L = F.lib()
    ks, km = ctypes.c_int(0), ctypes.c_int(0)
    L.fcfc_gpu_fastbin_scales(ns, nmu, periodic, ctypes.byref(ks), ctypes.byref(km))
    return ks.value, km.value


def exact_bins(d2, dz, ns, nmu, arith):
    """metric_common.c:170-184 (scalar order) / :472-494 (AVX-512 order), tables floor(sqrt(i)); -1 = rejected."""
    nmu2 = f32(nmu * nmu)
    sb = np.floor(np.sqrt(np.floor(d2.astype(np.float64)))).astype(np.int64)
    dz2 = dz * dz                                            # float32 product
    with np.errstate(divide="ignore", invalid="ignore"):
        if arith == 0:
            m = np.where(d2 < np.finfo(f32).eps, 0, np.trunc((dz2 / d2) * nmu2)).astype(np.int64)
        else:
            q = dz2 * nmu2
            qn = (q / d2).astype(f32)                        # round to nearest, then one ulp down if it overshot: RZ
            over = qn.astype(np.float64) * d2.astype(np.float64) > q.astype(np.float64)
            qz = np.where(over, np.nextafter(qn, f32(0)), qn)
            m = np.where(qz < nmu2, np.trunc(qz), nmu * nmu).astype(np.int64)
            m = np.where(d2 >= np.finfo(f32).eps, m, 0)
    mb = np.floor(np.sqrt(m.astype(np.float64))).astype(np.int64)
    ok = m < nmu * nmu
    return np.where(ok, sb, -1), np.where(ok, mb, -1)


def fast_bins(d2, dz, ns, nmu, ks, km, delta):
    r = (1.0 / np.sqrt(d2.astype(np.float64) + 1e-30) * (1.0 + delta)).astype(f32)
    sr, mr = d2 * r, np.abs(dz * r)                          # float32 products (FMUL2)
    i_s = np.floor(sr.astype(np.float64) * 2.0 ** ks).astype(np.int64) + 1        # fma.rz(sr, 2^ks, 2^23 + 1): exact
    i_m = np.floor(mr.astype(np.float64) * (nmu * 2.0 ** km)).astype(np.int64) + 1
    flagged = ((i_s & ((1 << ks) - 4)) == 0) | ((i_m & ((1 << km) - 2)) == 0)
    return i_s >> ks, i_m >> km, flagged


def sample_pairs(rng, n, ns, nmu):
    """(d2, dz) as float32, a third uniform, a third with s next to an integer, a third with nmu*mu next to an integer."""
    s = rng.uniform(0.05, ns, n)
    mu = rng.uniform(0, 1, n)
    k = n // 3
    near = rng.integers(1, ns + 1, k) + rng.normal(0, 1, k) * 2.0 ** rng.uniform(-26, -10, k)
    s[:k] = np.clip(near, 0.05, ns * (1 - 1e-7))
    nearm = (rng.integers(0, nmu + 1, k) + rng.normal(0, 1, k) * 2.0 ** rng.uniform(-26, -10, k)) / nmu
    mu[k:2 * k] = np.clip(nearm, 0, 1)
    d2 = (s * s).astype(f32)
    d2 = np.minimum(d2, np.nextafter(f32(ns * ns), f32(0)))  # the range test d2 < s2max passed
    dz = (np.sqrt(d2.astype(np.float64)) * mu * rng.choice([-1.0, 1.0], n)).astype(f32)
    dz = np.where(dz.astype(np.float64) ** 2 > d2, np.sign(dz) * np.sqrt(d2), dz).astype(f32)   # |dz| <= s as in a real pair
    return d2, dz


@pytest.mark.parametrize("ns,nmu", [(40, 120), (40, 1), (200, 50), (20, 255), (150, 100), (8, 10)])
@pytest.mark.parametrize("arith", [0, 1])
def test_unflagged_pairs_are_binned_exactly(ns, nmu, arith):
    ks, km = scales(ns, nmu)
    if ks < 6 or km < 6:
        pytest.skip("too many bins for the fixed-point trick: the engine uses the exact path")
    assert (ns + 2) * 2 ** ks < 2 ** 23 and (nmu + 2) * 2 ** km < 2 ** 23
    rng = np.random.default_rng(1000 * ns + nmu + arith)
    n = 600_000
    d2, dz = sample_pairs(rng, n, ns, nmu)
    es, em = exact_bins(d2, dz, ns, nmu, arith)
    nflag = 0
    for delta in (rng.uniform(-1, 1, n) * 2.0 ** -22, np.full(n, 2.0 ** -22), np.full(n, -2.0 ** -22)):
        fs, fm, flagged = fast_bins(d2, dz, ns, nmu, ks, km, delta)
        clean = ~flagged
        assert np.array_equal(fs[clean], es[clean]), "an unflagged pair got a different s bin"
        if nmu > 1:
            assert np.array_equal(fm[clean], em[clean]), "an unflagged pair got a different mu bin (or should have been rejected)"
        assert fs[clean].max() < ns and (nmu == 1 or fm[clean].max() < nmu)
        nflag += int(flagged.sum())
    # the band is narrow: on uniformly drawn pairs only a few per thousand take the exact path
    fs, fm, flagged = fast_bins(d2[2 * (n // 3):], dz[2 * (n // 3):], ns, nmu, ks, km, 0.0)
    assert flagged.mean() < 0.02


def test_scales_of_the_bench_workload():
    assert scales(40, 120) == (14, 13)


# ---------------------------------------------------------------------------------------------------
# Survey (s, mu): the queued pair is (t, s1, s2) with t = 2 x1.x2 and s_i = |x_i|^2 (2pt/metric_common.c:169-184);
# the drain rebuilds d2 = (s1 + s2) - t exactly as the pair loop did and takes pi = |s1 - s2| rsqrt(s1 + s2 + t) as the
# line-of-sight separation (count_kernel.cuh: fast_inputs), i.e. one more approximate reciprocal square root and, in
# double precision, three conversions to float; the host widens the mu band accordingly.
def survey_exact(t, s1, s2, ns, nmu, arith, real):
    nmu2 = real(nmu * nmu)
    s = s1 + s2
    d2 = s - t
    d = s1 - s2
    with np.errstate(divide="ignore", invalid="ignore"):
        num = (d * d) / (s + t)                               # pi^2
        if arith == 0:
            m = np.where(d2 < np.finfo(real).eps, 0, np.trunc((num / d2) * nmu2)).astype(np.int64)
        else:
            q = num * nmu2
            qn = (q / d2).astype(real)
            # round toward zero: one ulp down when the rounded quotient overshot (checked in higher precision)
            over = qn.astype(np.longdouble) * d2.astype(np.longdouble) > q.astype(np.longdouble)
            qz = np.where(over, np.nextafter(qn, real(0)), qn)
            m = np.where(qz < nmu2, np.trunc(qz), nmu * nmu).astype(np.int64)
            m = np.where(d2 >= np.finfo(real).eps, m, 0)
    sb = np.floor(np.sqrt(np.floor(np.maximum(d2, 0).astype(np.float64)))).astype(np.int64)
    mb = np.floor(np.sqrt(m.astype(np.float64))).astype(np.int64)
    ok = m < nmu * nmu
    return d2, np.where(ok, sb, -1), np.where(ok, mb, -1)


def survey_fast(t, s1, s2, ns, nmu, ks, km, delta1, delta2):
    s = s1 + s2
    d2f = np.maximum((s - t).astype(f32), f32(0))
    df, stf = (s1 - s2).astype(f32), (s + t).astype(f32)
    r1 = (1.0 / np.sqrt(stf.astype(np.float64)) * (1.0 + delta1)).astype(f32)
    auxf = df * r1
    r = (1.0 / np.sqrt(d2f.astype(np.float64) + 1e-30) * (1.0 + delta2)).astype(f32)
    sr, mr = d2f * r, np.minimum(np.abs(auxf * r), f32(1.000002))
    i_s = np.floor(sr.astype(np.float64) * 2.0 ** ks).astype(np.int64) + 1
    i_m = np.floor(mr.astype(np.float64) * (nmu * 2.0 ** km)).astype(np.int64) + 1
    flagged = ((i_s & ((1 << ks) - 4)) == 0) | ((i_m & ((1 << km) - 2)) == 0)
    return i_s >> ks, i_m >> km, flagged


@pytest.mark.parametrize("real", [np.float32, np.float64])
@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("ns,nmu", [(40, 100), (40, 120), (100, 30)])
def test_survey_unflagged_pairs_are_binned_exactly(ns, nmu, arith, real):
    ks, km = scales(ns, nmu, periodic=0)
    if ks < 6 or km < 6:
        pytest.skip("exact path")
    rng = np.random.default_rng(77 * ns + nmu + arith)
    n = 500_000
    # points at 200-340 from the observer in units of the s bin width (the rescaled survey), separations inside the range
    r1 = rng.uniform(200.0, 340.0, n)
    sep = rng.uniform(0.05, ns, n)
    k = n // 2
    sep[:k] = np.clip(rng.integers(1, ns + 1, k) + rng.normal(0, 1, k) * 2.0 ** rng.uniform(-24, -8, k), 0.05, ns * (1 - 1e-6))
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    x1 = (u * r1[:, None]).astype(real)
    x2 = x1.astype(np.float64) + v * sep[:, None]
    # a quarter of the pairs get nmu * mu planted next to an integer: mu is the cosine between the separation and the
    # direction of x1 + x2, which depends on x2 itself -- a few fixed-point iterations settle it
    q = n // 4
    mu_t = np.clip((rng.integers(0, nmu + 1, q) + rng.normal(0, 1, q) * 2.0 ** rng.uniform(-24, -8, q)) / nmu, 0, 1)
    perp = np.cross(u[-q:], v[-q:]); perp /= np.linalg.norm(perp, axis=1)[:, None]
    xa = x1[-q:].astype(np.float64)
    xb = x2[-q:].copy()
    for _ in range(10):
        h = xa + xb; h /= np.linalg.norm(h, axis=1)[:, None]
        e2 = perp - (perp * h).sum(1)[:, None] * h; e2 /= np.linalg.norm(e2, axis=1)[:, None]
        xb = xa + sep[-q:, None] * (mu_t[:, None] * h + np.sqrt(1 - mu_t ** 2)[:, None] * e2)
    x2[-q:] = xb
    x2 = x2.astype(real)
    s1 = ((x1[:, 0] * x1[:, 0] + x1[:, 1] * x1[:, 1]) + x1[:, 2] * x1[:, 2]).astype(real)
    s2 = ((x2[:, 0] * x2[:, 0] + x2[:, 1] * x2[:, 1]) + x2[:, 2] * x2[:, 2]).astype(real)
    t = (real(2) * ((x1[:, 0] * x2[:, 0] + x1[:, 1] * x2[:, 1]) + x1[:, 2] * x2[:, 2])).astype(real)
    d2, es, em = survey_exact(t, s1, s2, ns, nmu, arith, real)
    keep = (d2 < real(ns * ns)) & (d2 >= 0)                   # what the range test lets through
    uniform = ((np.arange(n) >= k) & (np.arange(n) < n - q))[keep]       # the pairs that were not planted next to a bin edge
    t, s1, s2, es, em = t[keep], s1[keep], s2[keep], es[keep], em[keep]
    m = len(t)
    e = 2.0 ** -22
    for d1, dd2 in ((rng.uniform(-e, e, m), rng.uniform(-e, e, m)), (e, e), (-e, -e), (e, -e), (-e, e)):
        fs, fm, flagged = survey_fast(t, s1, s2, ns, nmu, ks, km, d1, dd2)
        clean = ~flagged
        assert np.array_equal(fs[clean], es[clean]), "an unflagged pair got a different s bin"
        assert np.array_equal(fm[clean], em[clean]), "an unflagged pair got a different mu bin"
    assert flagged[uniform].mean() < 0.02                     # the band is narrow on ordinary pairs


# ---------------------------------------------------------------------------------------------------
# Survey (s_perp, pi): the pair loop keeps a candidate off the stacks with two division-free tests
# (count_kernel.cuh, eval_pair); the exact tests, with the division, run later (finish_pair).  Bit-exactness needs
# the cheap tests to be NECESSARY conditions of the exact ones through every rounding.
@pytest.mark.parametrize("real", [np.float32, np.float64])
@pytest.mark.parametrize("smax,pimax", [(40.0, 80.0), (10.0, 100.0), (150.0, 20.0)])
def test_survey_spi_pretests_never_drop_an_accepted_pair(real, smax, pimax):
    L = F.lib()
    lim = (ctypes.c_double * 3)()
    s2max, p2max = smax * smax, pimax * pimax                # the survey metric bins pi^2 (count_func.c:5296)
    L.fcfc_gpu_survey_pretest_limits(ctypes.c_double(s2max), ctypes.c_double(p2max), int(real is np.float32), lim)
    premax, pmax_pre, s2max_pre = (real(v) for v in lim)
    assert float(premax) >= s2max + p2max and float(pmax_pre) >= p2max and float(s2max_pre) >= s2max
    rng = np.random.default_rng(int(smax) * 7 + int(pimax))
    n = 1_500_000
    r1 = rng.uniform(200.0, 340.0, n)
    # separations in cylinder coordinates about the line of sight; two thirds planted on the cylinder's surfaces
    sperp = np.sqrt(rng.uniform(0, 1.2 * s2max, n))
    pi = rng.uniform(0, 1.2 * pimax, n)
    k = n // 3
    sperp[:k] = smax * (1 + rng.normal(0, 1, k) * 2.0 ** rng.uniform(-40, -12, k))
    pi[k:2 * k] = pimax * (1 + rng.normal(0, 1, k) * 2.0 ** rng.uniform(-40, -12, k))
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    w = rng.normal(size=(n, 3)); w -= (w * u).sum(1)[:, None] * u; w /= np.linalg.norm(w, axis=1)[:, None]
    xa = u * r1[:, None]
    xb = xa + pi[:, None] * u + sperp[:, None] * w
    for _ in range(8):                                       # the line of sight is along x1 + x2
        h = xa + xb; h /= np.linalg.norm(h, axis=1)[:, None]
        e2 = w - (w * h).sum(1)[:, None] * h; e2 /= np.linalg.norm(e2, axis=1)[:, None]
        xb = xa + pi[:, None] * h + sperp[:, None] * e2
    x1, x2 = xa.astype(real), xb.astype(real)
    s1 = ((x1[:, 0] * x1[:, 0] + x1[:, 1] * x1[:, 1]) + x1[:, 2] * x1[:, 2]).astype(real)
    s2 = ((x2[:, 0] * x2[:, 0] + x2[:, 1] * x2[:, 1]) + x2[:, 2] * x2[:, 2]).astype(real)
    t = (real(2) * ((x1[:, 0] * x2[:, 0] + x1[:, 1] * x2[:, 1]) + x1[:, 2] * x2[:, 2])).astype(real)
    # exact path (2pt/metric_common.c:169-205; finish_pair)
    s = s1 + s2
    d2 = s - t
    d = s1 - s2
    st = s + t
    dd = d * d
    num = dd / st
    accepted = (num < real(p2max)) & ((d2 - num) < real(s2max))
    # cheap path (eval_pair)
    ok = (d2 < premax) & (dd < st * pmax_pre) & ((d2 - s2max_pre) * st < dd)
    assert accepted.sum() > n // 10 and (~accepted).sum() > n // 10
    assert not np.any(accepted & ~ok), "a pre-test dropped a pair the exact tests accept"
    # and they are sharp: nearly everything they let through is accepted
    assert (ok & ~accepted).sum() < 0.02 * ok.sum() + 2 * k * 0.6


# ---------------------------------------------------------------------------------------------------
# Two small lemmas the exact re-binning path relies on (count_kernel.cuh: isqrt_small, div_rz_pos).
def test_isqrt_from_an_approximate_square_root_exhaustive():
    """floor(sqrt(m)) == trunc(sqrt.approx(m + 1/2)) for every integer 0 <= m < 2^18, whatever the sign of the
    approximation's error (2^-22 relative is far more than sqrt.approx.ftz.f32 is off)."""
    m = np.arange(1 << 18, dtype=np.int64)
    want = np.floor(np.sqrt(m.astype(np.float64))).astype(np.int64)
    arg = (m.astype(f32) + f32(0.5)).astype(np.float64)      # exact in float32 for m < 2^23
    for delta in (0.0, 2.0 ** -22, -(2.0 ** -22)):
        approx = (np.sqrt(arg) * (1.0 + delta)).astype(f32)
        assert np.array_equal(np.trunc(approx).astype(np.int64), want)


def test_round_toward_zero_quotient_from_the_rounded_one():
    """div_rz_pos: q = RN(a / b), one ulp down when the exact residual a - q b is negative, equals RZ(a / b)."""
    rng = np.random.default_rng(11)
    n = 2_000_000
    a = (rng.random(n) * 2.0 ** rng.integers(-10, 20, n)).astype(f32)
    b = (rng.random(n) * 2.0 ** rng.integers(-10, 20, n) + 1e-6).astype(f32)
    a[: n // 4] = (b[: n // 4].astype(np.float64) * rng.integers(1, 4000, n // 4)).astype(f32)      # (nearly) exact quotients
    q = (a / b).astype(f32)
    resid = a.astype(np.float64) - q.astype(np.float64) * b.astype(np.float64)      # 24 x 24 bits: exact in float64
    got = np.where(resid < 0, np.nextafter(q, f32(0)), q)
    # reference: the quotient in extended precision, truncated to 24 bits
    ql = a.astype(np.longdouble) / b.astype(np.longdouble)
    want = ql.astype(f32)
    want = np.where(want.astype(np.longdouble) > ql, np.nextafter(want, f32(0)), want)
    assert np.array_equal(got, want)
