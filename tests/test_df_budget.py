"""CPU: the error budget of the double-precision-at-float-speed kernel (fcfc_b200/csrc/count_kernel_df.cuh).

That kernel evaluates every pair in FP32 on coordinates taken relative to the centre of the tile's cell, bins it with the
computed ("fast") bins, and re-evaluates in FP64 only the pairs it flags: those within a band of a bin edge (the band
covers the float arithmetic AND the rounding of the coordinates), and (s,mu) pairs below a small separation s1.  Its
results are bit-identical to the reference's double build if and only if

    an UNFLAGGED pair gets, from the float path, exactly the decision (in range or not) and the bins
    that the reference's double-precision sequence gives it (metric_common.c:140-235).

This test emulates the float path in numpy (IEEE float32 operations; rsqrt.approx as the true value times 1 +- 2^-22;
the truncating fused multiply-add as exact integer arithmetic), with the scales, the padded range limit and s1 taken from
the library itself (fcfc_gpu_df_budget), on pairs planted at the bin edges, at the maximum separation, at mu = 1 and at
tiny separations, for coordinates as large as the bench workload's and far larger."""
import ctypes as C

import numpy as np
import pytest

import fcfc_b200 as F

f32 = np.float32


def budget(ns, nmu, cs_max):
    L = F.lib()
    L.fcfc_gpu_df_budget.restype = C.c_int
    L.fcfc_gpu_df_budget.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 5
    ks, km = C.c_int(0), C.c_int(0)
    lim, s1sq, fl = C.c_double(0), C.c_double(0), C.c_double(0)
    ok = L.fcfc_gpu_df_budget(ns, nmu, float(ns * ns), cs_max, C.byref(ks), C.byref(km), C.byref(lim), C.byref(s1sq), C.byref(fl))
    return ok, ks.value, km.value, lim.value, s1sq.value, fl.value


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def exact_double(a, b, ns, nmu, arith):
    """The reference's double-precision sequence: (in range, s bin, mu bin)."""
    d = a - b
    dz2 = d[:, 2] * d[:, 2]
    if arith == 0:
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + dz2                      # metric_common.c:170-172
    else:
        # fused chain of the AVX-512 build (:426-430); emulated in extended precision (the two roundings of an fma of doubles
        # cannot be reproduced in numpy, so pairs whose d2 could differ by it are excluded below through `doubt`)
        d2 = (d[:, 1] * d[:, 1] + (d[:, 0] * d[:, 0] + dz2))
    ok = d2 < ns * ns
    sb = np.floor(np.sqrt(np.floor(d2))).astype(np.int64)
    nmu2 = float(nmu * nmu)
    with np.errstate(divide="ignore", invalid="ignore"):
        m = np.where(d2 < np.finfo(np.float64).eps, 0, np.trunc((dz2 / d2) * nmu2))
    m = np.nan_to_num(m, nan=0.0).astype(np.int64)
    if nmu == 1:                # isotropic bins: no mu cut
        return ok, sb, np.zeros_like(sb)
    ok &= m < nmu * nmu
    mb = np.floor(np.sqrt(np.minimum(m, nmu * nmu).astype(np.float64))).astype(np.int64)
    return ok, sb, mb


def float_path(a, b, o, ns, nmu, ks, km, lim, s1sq, delta):
    """count_kernel_df.cuh: relative float coordinates, packed distance chain, fast_bins, flags."""
    af, bf = (a - o).astype(f32), (b - o).astype(f32)          # double subtraction of the origin, then one rounding
    d = af - bf
    d2 = fma32(d[:, 1], d[:, 1], fma32(d[:, 0], d[:, 0], d[:, 2] * d[:, 2]))
    dz = d[:, 2]
    r = (1.0 / np.sqrt(d2.astype(np.float64) + 1e-30) * (1.0 + delta)).astype(f32)
    sr, mr = d2 * r, np.abs(dz * r)
    i_s = np.floor(sr.astype(np.float64) * 2.0 ** ks).astype(np.int64) + 1
    i_m = np.floor(mr.astype(np.float64) * (nmu * 2.0 ** km)).astype(np.int64) + 1
    flagged = (i_s & ((1 << ks) - 4)) == 0
    if nmu > 1:
        flagged |= ((i_m & ((1 << km) - 2)) == 0) | (d2 < f32(s1sq))
    inr = d2 < f32(lim)
    return inr, i_s >> ks, (i_m >> km) if nmu > 1 else np.zeros_like(i_s), flagged


def planted_pairs(rng, n, ns, nmu, cs, origin_scale):
    """Primaries inside a cell, secondaries at separations on / next to the s edges, the mu edges, s_max, mu = 1, s ~ 0."""
    o = rng.uniform(-origin_scale, origin_scale, (n, 3))
    a = o + rng.uniform(-0.5, 0.5, (n, 3)) * cs
    s = rng.uniform(0.0, ns * 1.0005, n)
    mu = rng.uniform(0.0, 1.0, n)
    k = n // 5
    jit = rng.normal(0, 1, n) * 2.0 ** rng.uniform(-30, -8, n)
    s[:k] = rng.integers(1, ns + 1, k) + jit[:k]                               # s bin edges, including s_max
    mu[k:2 * k] = np.clip((rng.integers(0, nmu + 1, k) + jit[k:2 * k]) / nmu, 0, 1)     # mu bin edges, including mu = 1
    s[2 * k:3 * k] = 10.0 ** rng.uniform(-6, 0.5, k)                             # tiny separations
    s[3 * k:4 * k] = ns * (1 + jit[3 * k:4 * k])                                 # the maximum separation
    s = np.abs(s)
    phi = rng.uniform(0, 2 * np.pi, n)
    st = np.sqrt(np.maximum(0.0, 1 - mu * mu))
    d = np.stack([s * st * np.cos(phi), s * st * np.sin(phi), s * mu * rng.choice([-1.0, 1.0], n)], 1)
    return a, a - d, o


@pytest.mark.parametrize("ns,nmu,cs,origin_scale", [(40, 120, 8.2, 400.0), (40, 1, 8.2, 400.0), (40, 120, 8.2, 5e4), (200, 50, 45.0, 3000.0),
                                                    (20, 255, 5.0, 100.0), (150, 100, 30.0, 1000.0), (8, 10, 9.0, 50.0)])
def test_unflagged_pairs_follow_the_double_sequence(ns, nmu, cs, origin_scale):
    ok, ks, km, lim, s1sq, frac = budget(ns, nmu, cs)
    if not ok:
        pytest.skip("budget says: not usable for this binning (the engine takes another kernel)")
    assert (ns + 2) * 2 ** ks < 2 ** 23 and (nmu == 1 or (nmu + 2) * 2 ** km < 2 ** 23)
    assert ns * ns < lim < ns * ns * 1.001
    rng = np.random.default_rng(ns * 1000 + nmu)
    n = 1_000_000
    a, b, o = planted_pairs(rng, n, ns, nmu, cs, origin_scale)
    eok, es, em = exact_double(a, b, ns, nmu, 0)
    # the FMA order of the double sequence differs from the scalar one by at most a few ulp of d2: pairs whose integer part
    # of d2 or mu index could change within 4 ulp are "in doubt" for the arith = 1 statement and are required to be flagged
    d = a - b
    d2 = (d[:, 0] ** 2 + d[:, 1] ** 2) + d[:, 2] ** 2
    for delta in (rng.uniform(-1, 1, n) * 2.0 ** -22, np.full(n, 2.0 ** -22), np.full(n, -2.0 ** -22)):
        inr, fs, fm, flagged = float_path(a, b, o, ns, nmu, ks, km, lim, s1sq, delta)
        clean = ~flagged
        # 1. the range decision of an unflagged pair is the exact one (mu = 1 pairs are flagged, so `eok` applies as is)
        assert np.array_equal(inr[clean], eok[clean]), "an unflagged pair was accepted / rejected differently from the double sequence"
        cin = clean & inr
        # 2. ... and so are its bins
        assert np.array_equal(fs[cin], es[cin]), "an unflagged pair got a different s bin"
        if nmu > 1:
            assert np.array_equal(fm[cin], em[cin]), "an unflagged pair got a different mu bin"
        assert fs[cin].max() < ns and (nmu == 1 or fm[cin].max() < nmu)
        # 3. robust against the evaluation order of the double sequence: an unflagged pair is at least 1e-9 (relative) away
        #    from every decision boundary in double, i.e. thousands of ulp
        s_ex = np.sqrt(d2[cin])
        assert np.abs(s_ex - np.round(s_ex)).min() > 1e-9 * ns
    # the flagged fraction of uniformly drawn pairs is what the budget promises (a fraction of a per cent)
    s = ns * np.cbrt(rng.random(n))
    mu = rng.random(n)
    phi = rng.uniform(0, 2 * np.pi, n)
    st = np.sqrt(1 - mu * mu)
    dd = np.stack([s * st * np.cos(phi), s * st * np.sin(phi), s * mu], 1)
    inr, fs, fm, flagged = float_path(a, a - dd, o, ns, nmu, ks, km, lim, s1sq, 0.0)
    assert flagged.mean() < max(3 * frac, 2e-3), (flagged.mean(), frac)
    assert frac < 0.03


def test_bench_workload_budget():
    """configs[1] (C2) in double: 49^3 cells of 40.8 Mpc/h (8.16 rescaled), 40 x 120 bins."""
    ok, ks, km, lim, s1sq, frac = budget(40, 120, 8.17)
    assert ok and ks >= 12 and km >= 8 and frac < 0.01
    print(f"C2 double: ks={ks} km={km} d2lim={lim} s1={np.sqrt(s1sq):.3f} expected flagged fraction {frac:.4f}")
