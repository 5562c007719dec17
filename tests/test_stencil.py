"""CPU: the static neighbour-cell stencil that replaces the reference's dual-tree pruning (compare_dual_node,
src/fcfc/2pt_box/metric_kdtree.c:50-289) and its dense sub-ranges, as the counting path builds them
(engine.cu: build_stencil, stencil_inside; exported for this test as fcfc_gpu_debug_stencil).
  * complete: the cell offset of every pair closer than the reach is in the stencil (half stencil: it or its mirror);
  * no waste: every listed cell can hold such a pair (its nearest corner is within reach);
  * dense cells really are dense: the farthest corners of the two cells are closer than the maximum separation, and
    the classification is maximal (the next cell along z is not)."""
import ctypes

import numpy as np
import pytest

import fcfc_b200 as F


def stencil(cs, r2, s2max, half):
    L = F.lib()
    n = 4096
    rows, ins = (ctypes.c_int * (4 * n))(), (ctypes.c_int * (2 * n))()
    k = L.fcfc_gpu_debug_stencil((ctypes.c_double * 3)(*cs), ctypes.c_double(r2), ctypes.c_double(s2max), int(half), rows, ins, n)
    assert 0 < k <= n
    return np.array(rows[: 4 * k]).reshape(k, 4), np.array(ins[: 2 * k]).reshape(k, 2)


CASES = [((40.0, 40.0, 40.0), 200.0), ((40.8, 40.8, 40.8), 200.0), ((14.3, 14.3, 14.3), 42.0), ((10.0, 13.0, 7.0), 30.0),
         ((55.0, 50.0, 100.0), 100.0), ((3.0, 3.0, 3.0), 40.0)]


@pytest.mark.parametrize("cs,smax", CASES)
@pytest.mark.parametrize("half", [0, 1])
def test_stencil_is_complete_and_tight(cs, smax, half):
    cs = np.array(cs)
    s2max = smax * smax
    r2 = s2max * (1 + 1e-6)
    rows, _ = stencil(cs, r2, s2max, half)
    cells = {(int(r[0]), int(r[1]), z) for r in rows for z in range(int(r[2]), int(r[3]) + 1)}
    assert len(cells) == sum(int(r[3] - r[2] + 1) for r in rows), "rows overlap"
    if half:
        assert (0, 0, 0) not in cells and not any((-a, -b, -c) in cells for (a, b, c) in cells), "half stencil lists a pair twice"
    # complete: random pairs closer than smax
    rng = np.random.default_rng(3)
    n = 400_000
    p = rng.random((n, 3)) * cs
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    q = p + v * (smax * rng.random(n) ** (1 / 3))[:, None]
    off = np.floor(q / cs).astype(int)
    for o in {tuple(x) for x in off}:
        if o == (0, 0, 0):
            continue
        assert o in cells or (half and tuple(-c for c in o) in cells), f"offset {o} holds a pair in range but is not swept"
    # tight: the nearest corners of cell 0 and every listed cell are within reach
    for (a, b, c) in cells:
        gap = np.array([max(abs(a) - 1, 0), max(abs(b) - 1, 0), max(abs(c) - 1, 0)]) * cs
        assert (gap ** 2).sum() < r2, f"offset {(a, b, c)} cannot hold a pair in range"


@pytest.mark.parametrize("cs,smax", CASES)
def test_dense_cells_are_entirely_in_range(cs, smax):
    cs = np.array(cs)
    s2max = smax * smax
    rows, ins = stencil(cs, s2max * (1 + 1e-6), s2max, 0)
    ndense = 0
    for (dx, dy, zlo, zhi), (ilo, ihi) in zip(rows, ins):
        far_xy = ((abs(dx) + 1) * cs[0]) ** 2 + ((abs(dy) + 1) * cs[1]) ** 2
        if ilo > ihi:
            # no dense cell in this row: not even the nearest cell along z (dz = 0, or the row's closest one) qualifies
            zmin = min(range(zlo, zhi + 1), key=abs)
            assert far_xy + ((abs(zmin) + 1) * cs[2]) ** 2 >= s2max * (1 - 1e-4)
            continue
        assert zlo <= ilo <= ihi <= zhi
        for dz in range(ilo, ihi + 1):
            ndense += 1
            assert far_xy + ((abs(dz) + 1) * cs[2]) ** 2 < s2max, "a dense cell has a corner pair out of range"
        for dz in (ilo - 1, ihi + 1):                        # maximal: the next cells along z are not dense
            if zlo <= dz <= zhi:
                assert far_xy + ((abs(dz) + 1) * cs[2]) ** 2 >= s2max * (1 - 1e-4)
    if smax / cs.max() >= 3:
        assert ndense > 0


def test_bench_workload_stencil():
    """C2: cells of 2000/49, reach 200: 171 dense cells of 1015 swept."""
    c = 2000.0 / 49
    rows, ins = stencil((c, c, c), 200.0 ** 2 * (1 + 1e-6), 200.0 ** 2, 0)
    swept = sum(int(r[3] - r[2] + 1) for r in rows)
    dense = sum(int(i[1] - i[0] + 1) for i in ins if i[0] <= i[1])
    assert (swept, dense) == (1015, 171), (swept, dense)
