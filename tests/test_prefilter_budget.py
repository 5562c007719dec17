"""CPU: the single-precision pre-filter of the double-precision kernels (fcfc_b200/csrc/count_kernel_pf.cuh) never drops
a pair that the exact double-precision tests accept.

The filter's arithmetic is emulated in numpy float32 (packed f32x2 operations are ordinary IEEE float operations) with
the padded limits the library itself computes (fcfc_gpu_prefilter_limits); the exact tests follow the reference's
formulas in double (box: metric_common.c:140-235; survey: 2pt/metric_common.c:169-205).  Pairs are planted on the
surfaces of the accepted regions, where a filter with too little padding would fail first, at the coordinate
magnitudes of the bench workloads and far beyond."""
import ctypes as C

import numpy as np
import pytest

import fcfc_b200 as F

f32 = np.float32


def limits(periodic, bintype, s2max, pmax, M, smax_sq=0.0, smin_sq=0.0):
    L = F.lib()
    L.fcfc_gpu_prefilter_limits.restype = C.c_int
    L.fcfc_gpu_prefilter_limits.argtypes = [C.c_int, C.c_int] + [C.c_double] * 5 + [C.POINTER(C.c_double)]
    out = (C.c_double * 4)()
    mode = L.fcfc_gpu_prefilter_limits(int(periodic), bintype, s2max, pmax, M, smax_sq, smin_sq, out)
    return mode, [float(v) for v in out]


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def directions(n, rng):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


@pytest.mark.parametrize("M,rmax", [(800.0, 40.0), (4000.0, 40.0), (300.0, 80.0), (1e5, 25.0)])
def test_sphere_filter_is_a_superset(M, rmax):
    """Box / isotropic / (s,mu): d^2 < s2max exactly  ==>  float d^2 < padded limit."""
    rng = np.random.default_rng(1)
    s2max = rmax * rmax
    mode, lim = limits(1, 1, s2max, 0.0, M)
    assert mode == 1 and lim[3] < 0.05
    n = 400000
    a = rng.uniform(-M, M, (n, 3))
    r = rmax * (1 + rng.uniform(-3e-6, 1e-7, n))              # on and just inside the sphere
    b = a + directions(n, rng) * r[:, None]
    keep = np.abs(b).max(axis=1) <= M
    a, b = a[keep], b[keep]
    d = a - b
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    exact = d2 < s2max
    assert exact.sum() > 1000
    df = a.astype(f32) - b.astype(f32)
    d2f = fma32(df[:, 2], df[:, 2], fma32(df[:, 1], df[:, 1], df[:, 0] * df[:, 0]))
    passed = d2f < f32(lim[0])
    assert not (exact & ~passed).any()
    # and the padding is small: the filter passes few pairs the exact test rejects
    assert lim[0] / s2max - 1 < 0.05


def test_box_cylinder_filter_is_a_superset():
    rng = np.random.default_rng(2)
    M, smax, pmax = 800.0, 20.0, 24.0
    mode, lim = limits(1, 2, smax * smax, pmax, M)
    assert mode == 1
    n = 400000
    a = rng.uniform(-M + 50, M - 50, (n, 3))
    phi = rng.uniform(0, 2 * np.pi, n)
    on_side = rng.random(n) < 0.5
    rho = np.where(on_side, smax * (1 + rng.uniform(-3e-6, 1e-7, n)), smax * np.sqrt(rng.random(n)))
    z = np.where(on_side, pmax * rng.uniform(-1, 1, n), pmax * (1 + rng.uniform(-3e-6, 1e-7, n)) * rng.choice([-1, 1], n))
    b = a + np.stack([rho * np.cos(phi), rho * np.sin(phi), z], 1)
    d = a - b
    exact = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) < smax * smax) & (np.abs(d[:, 2]) < pmax)
    assert exact.sum() > 1000
    df = a.astype(f32) - b.astype(f32)
    passed = (fma32(df[:, 1], df[:, 1], df[:, 0] * df[:, 0]) < f32(lim[0])) & (np.abs(df[:, 2]) < f32(lim[1]))
    assert not (exact & ~passed).any()


@pytest.mark.parametrize("rlo,rhi,smax,pmax", [(500.0, 850.0, 20.0, 40.0), (150.0, 400.0, 20.0, 40.0), (2000.0, 9000.0, 30.0, 30.0)])
def test_survey_cylinder_filter_is_a_superset(rlo, rhi, smax, pmax):
    """Survey (s_perp, pi): the exact tests pi^2 < p2max, s_perp^2 = s^2 - pi^2 < s2max (2pt/metric_common.c:180-205) imply the
    three float tests of the filter (sphere, dd < st * plim, (d2 - s2lim) st <= dd)."""
    rng = np.random.default_rng(3)
    s2max, p2max = smax * smax, pmax * pmax
    mode, lim = limits(0, 2, s2max, p2max, rhi, rhi * rhi, rlo * rlo)
    assert mode == 2, "the cylinder tests should be usable at these distances"
    n = 600000
    u = directions(n, rng)
    a = u * rng.uniform(rlo, rhi, n)[:, None]
    # offsets in the frame of the line of sight: on the side surface, on the caps, and inside
    kind = rng.integers(0, 3, n)
    rho = np.where(kind == 0, smax * (1 + rng.uniform(-3e-6, 1e-7, n)), smax * np.sqrt(rng.random(n)))
    zz = np.where(kind == 1, pmax * (1 + rng.uniform(-3e-6, 1e-7, n)) * rng.choice([-1, 1], n), pmax * rng.uniform(-1, 1, n))
    t1 = np.cross(u, directions(n, rng)); t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    b = a + u * zz[:, None] + t1 * rho[:, None]
    keep = (np.linalg.norm(b, axis=1) >= rlo) & (np.linalg.norm(b, axis=1) <= rhi)
    a, b = a[keep], b[keep]
    s1 = (a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1]) + a[:, 2] * a[:, 2]
    s2 = (b[:, 0] * b[:, 0] + b[:, 1] * b[:, 1]) + b[:, 2] * b[:, 2]
    t = 2 * ((a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2])
    s = s1 + s2
    d2 = s - t
    ds = s1 - s2
    pi2 = ds * ds / (s + t)
    exact = (pi2 < p2max) & (d2 - pi2 < s2max)
    assert exact.sum() > 1000 and (~exact).sum() > 1000
    af, bf, s1f, s2f = a.astype(f32), b.astype(f32), s1.astype(f32), s2.astype(f32)
    df = af - bf
    d2f = fma32(df[:, 2], df[:, 2], fma32(df[:, 1], df[:, 1], df[:, 0] * df[:, 0]))
    sdif, ssum = s1f - s2f, s1f + s2f
    st = (ssum + ssum) - d2f
    dd = sdif * sdif
    passed = (d2f < f32(lim[0])) & (dd < st * f32(lim[1])) & ((d2f - f32(lim[2])) * st <= dd)
    assert not (exact & ~passed).any()
    # the cylinder tests are worth having: most of the sphere's candidates outside the cylinder are rejected
    sphere_only = d2f < f32(lim[0])
    assert (passed & ~exact).sum() < 0.2 * (sphere_only & ~exact).sum() + 50


def test_filter_switches_itself_off():
    """Tiny separations at huge coordinates: the padding would exceed 5 % -> the plain double kernel is used.
    Observer inside the sample: no cylinder tests, sphere only."""
    mode, lim = limits(1, 0, 1.0, 0.0, 1e6)
    assert mode == 0
    mode, lim = limits(0, 2, 400.0, 1600.0, 850.0, 850.0 ** 2, 1.0)
    assert mode == 1 and lim[1] == 0.0 and lim[2] == 0.0
