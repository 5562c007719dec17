"""CPU: the classification of staged secondary points against the tile (count_kernel_cl.cuh, count_kernel_pf.cuh) never drops
a point that has a partner in range, and never calls a point "dense" that has a partner out of range.

The counting kernels test every secondary point they stage against the bounding box of the tile's primaries and drop it when
no pair with a point of the box can be accepted.  Exactness rests on that test being a superset test THROUGH the float
arithmetic it is made in; this file emulates the device arithmetic in numpy float32 on boxes and points planted at the limits:

  * box / isotropic counts, single precision (count_kernel_cl.cuh): squared distance to the nearest point of the box against
    the padded limit of fcfc_gpu_classify_limits, squared distance to the farthest corner against the dense limit; the pair
    arithmetic (FMA order of metric_common.c:426-430 and scalar order :170-172) decides what "in range" means;
  * survey (s_perp, pi) counts (count_kernel_pf.cuh): the sphere test and the two cylinder bounds
        s_perp^2 = 4 |a x b|^2 / |a + b|^2 >= 2 |b|^2 (rho - R)^2 / (max|a|^2 + |b|^2)
        pi^2     = (|a|^2 - |b|^2)^2 / |a + b|^2 >= gap^2 / (2 (max|a|^2 + |b|^2))
    against the exact double-precision tests of 2pt/metric_common.c:169-205."""
import ctypes as C

import numpy as np
import pytest

import fcfc_b200 as F

f32 = np.float32


def classify_limits(s2max, maxabs):
    L = F.lib()
    L.fcfc_gpu_classify_limits.restype = None
    L.fcfc_gpu_classify_limits.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_float)]
    out = (C.c_float * 2)()
    L.fcfc_gpu_classify_limits(s2max, maxabs, out)
    return f32(out[0]), f32(out[1])


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def pair_d2_float(a, b, arith):
    """d^2 of the float kernels: FMA order (metric_common.c:426-430) or scalar order (:170-172)."""
    dx, dy, dz = (a[..., 0] - b[..., 0]).astype(f32), (a[..., 1] - b[..., 1]).astype(f32), (a[..., 2] - b[..., 2]).astype(f32)
    if arith:
        return fma32(dy, dy, fma32(dx, dx, (dz * dz).astype(f32)))
    return ((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32) + (dz * dz).astype(f32)


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("L,rmax,shift", [(2000.0, 200.0, 0.0), (2000.0, 200.0, 2000.0), (1.0e5, 50.0, 1.0e5), (300.0, 40.0, 300.0)])
def test_box_classification_is_safe(arith, L, rmax, shift):
    """Boxes of primaries anywhere in [0, L) (+ an image shift), secondaries planted on both sides of the drop limit and of the
    dense limit: a dropped point has no partner with float d^2 < s2max, a dense point only partners in range."""
    rng = np.random.default_rng(11)
    s2max = f32(rmax * rmax)
    maxabs = max(L, 2 * L if shift else L)
    skip, dense = classify_limits(float(s2max), maxabs)
    ndrop = ndense = 0
    for _ in range(300):
        lo = rng.random(3) * (L - rmax * 0.4)
        ext = rng.random(3) * rmax * 0.35
        prim = (lo + rng.random((64, 3)) * ext).astype(f32)
        prim[:8] = (lo + rng.integers(0, 2, (8, 3)) * ext).astype(f32)             # corners
        sh = f32(shift) if rng.random() < 0.5 else f32(0)
        a = (prim + sh).astype(f32)                                                 # the image-shifted primaries of the pair loop
        # device bounding box: of the unshifted tile, centre shifted afterwards (count_kernel_cl.cuh)
        blo, bhi = prim.min(0), prim.max(0)
        c = (f32(0.5) * (blo + bhi)).astype(f32)
        h = (f32(0.5) * (bhi - blo)).astype(f32)
        c = (c + sh).astype(f32)
        # secondaries: directions from the box, distances hugging rmax (+- a few 1e-5 relative) and farther out / closer in
        n = 4000
        u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
        corner = c.astype(np.float64) + np.sign(u) * h.astype(np.float64)
        scale = rmax * (1 + np.concatenate([rng.normal(0, 3e-5, n // 2), rng.normal(0, 0.2, n - n // 2)]))
        far = rng.random(n) < 0.5                                                   # half aimed at the nearest corner, half at the farthest
        b = np.where(far[:, None], c.astype(np.float64) - np.sign(u) * h.astype(np.float64) + u * scale[:, None], corner + u * scale[:, None]).astype(f32)
        t = np.abs((b - c).astype(f32))
        un = np.maximum((t - h).astype(f32), f32(0))
        dmin2 = fma32(un[:, 2], un[:, 2], fma32(un[:, 1], un[:, 1], (un[:, 0] * un[:, 0]).astype(f32)))
        v = (t + h).astype(f32)
        dmax2 = fma32(v[:, 2], v[:, 2], fma32(v[:, 1], v[:, 1], (v[:, 0] * v[:, 0]).astype(f32)))
        drop = dmin2 > skip
        isdense = ~drop & (dmax2 < dense)
        d2 = pair_d2_float(a[None, :, :], b[:, None, :], arith)                    # [secondary, primary]
        inr = d2 < s2max
        assert not inr[drop].any(), "a dropped point has a partner in range"
        assert inr[isdense].all(), "a dense point has a partner out of range"
        ndrop += int(drop.sum()); ndense += int(isdense.sum())
    assert ndrop > 1000 and ndense > 1000                                           # the limits were actually probed


def test_survey_cylinder_classification_is_safe():
    """Survey (s_perp, pi): tiles at 800-3000 from the observer, secondaries planted around the cylinder's surfaces; a dropped
    point must fail the exact double-precision tests (pi^2 < p2max and s_perp^2 = s^2 - pi^2 < s2max) with every primary."""
    rng = np.random.default_rng(12)
    ndrop = nacc = 0
    for it in range(400):
        smax, pmax = rng.choice([20.0, 40.0, 80.0]), rng.choice([40.0, 80.0, 120.0])
        s2max, p2max = smax * smax, pmax * pmax
        dist = rng.uniform(800, 3000)
        dirn = rng.normal(size=3); dirn /= np.linalg.norm(dirn)
        ext = rng.uniform(2, 30, 3)
        prim = dist * dirn + (rng.random((48, 3)) - 0.5) * ext
        prim[:8] = dist * dirn + (rng.integers(0, 2, (8, 3)) - 0.5) * ext
        pf = prim.astype(f32)
        ps = (prim ** 2).sum(1)
        lo, hi = pf.min(0), pf.max(0)
        h = (f32(0.5) * (hi - lo)).astype(f32); c = (f32(0.5) * (lo + hi)).astype(f32)
        R = f32(np.sqrt(f32(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]))) * f32(1.001) + f32(1e-6) * np.maximum(np.abs(lo), np.abs(hi)).max().astype(f32)
        smn, smx = f32(ps.min()) * f32(0.999999), f32(ps.max()) * f32(1.000001)
        f_d2lim = f32((s2max + p2max) * (1 + 1e-5))                                  # (the filter's padded sphere, never below the exact one)
        f_s2cl, f_p2cl = f32(s2max) * f32(1.001), f32(p2max) * f32(1.001)
        # secondaries: offsets from random primaries with (s_perp, pi) hugging the limits, and a broad cloud
        n = 3000
        base = prim[rng.integers(0, len(prim), n)]
        los = base / np.linalg.norm(base, axis=1)[:, None]
        w = rng.normal(size=(n, 3)); w -= (w * los).sum(1)[:, None] * los; w /= np.linalg.norm(w, axis=1)[:, None]
        sp = np.where(rng.random(n) < 0.5, smax * (1 + rng.normal(0, 0.02, n)), rng.uniform(0, 2.5 * smax, n))
        pi = np.where(rng.random(n) < 0.5, pmax * (1 + rng.normal(0, 0.02, n)), rng.uniform(0, 2.0 * pmax, n)) * rng.choice([-1, 1], n)
        b = base + w * sp[:, None] + los * pi[:, None]
        bs = (b ** 2).sum(1)
        bf, bsf = b.astype(f32), bs.astype(f32)
        # device classification (count_kernel_pf.cuh), float32
        vv = (bf - c).astype(f32)
        un = np.maximum((np.abs(vv) - h).astype(f32), f32(0))
        keep = ~(fma32(un[:, 2], un[:, 2], fma32(un[:, 1], un[:, 1], (un[:, 0] * un[:, 0]).astype(f32))) > f_d2lim * f32(1.001))
        kx = (c[1] * bf[:, 2]).astype(f32) - (c[2] * bf[:, 1]).astype(f32)
        ky = (c[2] * bf[:, 0]).astype(f32) - (c[0] * bf[:, 2]).astype(f32)
        kz = (c[0] * bf[:, 1]).astype(f32) - (c[1] * bf[:, 0]).astype(f32)
        cross2 = fma32(kz, kz, fma32(ky, ky, (kx * kx).astype(f32)))
        ssum = (smx + bsf).astype(f32)
        tt = (np.sqrt((f_s2cl * ssum / (f32(2) * bsf)).astype(f32)).astype(f32) * f32(1.001) + R).astype(f32)
        keep &= ~(cross2 > (tt * tt * bsf).astype(f32) * f32(1.001))
        gap = np.maximum(np.maximum((bsf - smx).astype(f32), (smn - bsf).astype(f32)), f32(0))
        keep &= ~((gap * gap).astype(f32) > (f_p2cl * f32(2) * ssum).astype(f32))
        # exact tests of the reference for every (secondary, primary): 2pt/metric_common.c:169-205
        t2 = 2 * (b[:, None, :] * prim[None, :, :]).sum(-1)
        s = bs[:, None] + ps[None, :]
        d = ps[None, :] - bs[:, None]
        pi2 = d * d / (s + t2)
        sperp2 = (s - t2) - pi2
        acc = (pi2 < p2max) & (sperp2 < s2max)
        assert not acc[~keep].any(), "a dropped point has an accepted partner"
        ndrop += int((~keep).sum()); nacc += int(acc.any(1).sum())
    assert ndrop > 50000 and nacc > 50000
