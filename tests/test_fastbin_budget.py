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
