"""GPU (-m gpu): streamed ingest (include/fcfc_gpu.h: fcfc_gpu_catalog_stream_*) -- SURVEY section 8(f) rank 1.

A catalogue appended chunk by chunk, the way the reference's reader produces it (io/read_ascii.c:750-950: fread a chunk of
the file, parse its lines, append the rows to the columns), must be THE SAME catalogue as the one-shot upload of the
concatenated columns: identical counts bit for bit (weighted sums within the 1e-12 of north_star: the order of the
atomic additions is not fixed), the same sum of weights, the same errors.  Edge cases: empty chunks, chunks larger than a
pinned staging slot, growth past the size hint, no hint at all, an empty catalogue, reuse of the caller's buffer right
after append() returns."""
import io

import numpy as np
import pytest

from cases import box_catalog, survey_catalog

pytestmark = pytest.mark.gpu

WT_RTOL = 1e-12        # weighted FP64 sums depend on the order of the atomic additions at the 1e-16 level (north_star: 1e-12)


def same(got, want, withwt):
    if withwt:
        np.testing.assert_array_equal(got == 0, want == 0)
        np.testing.assert_allclose(got, want, rtol=WT_RTOL, atol=0)
    else:
        np.testing.assert_array_equal(got, want)


BOX_KW = dict(box=500.0, bintype=1, smin=0.0, smax=50.0, ds=2.5, nmu=20)


def ragged_chunks(n, seed, big=None):
    """Chunk boundaries with empty and tiny chunks, optionally one chunk of `big` rows."""
    rng = np.random.default_rng(seed)
    cuts, pos = [0], 0
    if big:
        pos = min(n, big)
        cuts.append(pos)
    while pos < n:
        step = int(rng.choice([0, 1, 7, 1000, 4096, 33333]))
        pos = min(n, pos + step)
        cuts.append(pos)
    return list(zip(cuts[:-1], cuts[1:]))


def stream_catalog(F, cols, bins, chunks, n_hint=0, weighted=False, reuse_buffer=False):
    st = F.CatalogStream(bins=bins, weighted=weighted, n_hint=n_hint)
    scratch = [np.empty(max((e - b for b, e in chunks), default=0) or 1, dtype=bins.dtype) for _ in cols]
    for b, e in chunks:
        if reuse_buffer:
            # the caller's arrays are free again when append() returns: overwrite them at once
            views = []
            for s, c in zip(scratch, cols):
                s[: e - b] = c[b:e]
                views.append(s[: e - b])
            st.append(*views[:3], views[3] if weighted else None)
            for s in scratch:
                s[:] = np.nan
        else:
            st.append(*(c[b:e] for c in cols[:3]), cols[3][b:e] if weighted else None)
    assert len(st) == len(cols[0])
    return st.finish()


@pytest.mark.parametrize("prec", ["float", "double"])
@pytest.mark.parametrize("withwt", [False, True])
def test_streamed_box_catalogue_equals_one_shot(gpu, prec, withwt):
    F = gpu
    n = 300_000                 # more than the four pinned slots together (2^16 rows each)
    cols = box_catalog(n, 500.0, 91, weights=True)
    b = F.Bins(periodic=True, prec=prec, arith=1, **BOX_KW)
    one = F.Catalog(*cols[:3], cols[3] if withwt else None, bins=b)
    want = F.count_pairs(one, None, b, withwt=withwt)
    assert want.sum() > 0
    for kw in (dict(n_hint=n), dict(n_hint=0), dict(n_hint=1000, reuse_buffer=True)):
        chunks = ragged_chunks(n, 5, big=280_000 if kw.get("n_hint") == n else None)
        cat = stream_catalog(F, cols, b, chunks, weighted=withwt, **kw)
        assert cat.n == n
        got = F.count_pairs(cat, None, b, withwt=withwt)
        same(got, want, withwt)
        assert cat.wsum == pytest.approx(one.wsum, rel=1e-13)
        # and as the secondary of a cross count with the one-shot catalogue
        same(F.count_pairs(one, cat, b, withwt=withwt), F.count_pairs(cat, one, b, withwt=withwt), withwt)
        cat.destroy()
    one.destroy()


def test_streamed_survey_catalogue_computes_the_fourth_coordinate(gpu):
    """Survey (s,mu) needs x^2+y^2+z^2 per point (2pt/build_tree.c:59 / :75-82): finish() forms it on the device in the
    order of the bins, as the one-shot upload does."""
    F = gpu
    d, r = survey_catalog(20000, 5), survey_catalog(50000, 6)
    for arith in (0, 1):
        b = F.Bins(periodic=False, prec="double", bintype=1, smin=0.0, smax=120.0, ds=4.0, nmu=50, arith=arith)
        D1, R1 = F.Catalog(*d, bins=b), F.Catalog(*r, bins=b)
        D2 = stream_catalog(F, d, b, ragged_chunks(len(d[0]), 1), weighted=True)
        R2 = stream_catalog(F, r, b, ragged_chunks(len(r[0]), 2), n_hint=len(r[0]), weighted=True)
        for p, q, pp, qq in ((D1, None, D2, None), (D1, R1, D2, R2), (R1, None, R2, None)):
            same(F.count_pairs(pp, qq, b, withwt=True), F.count_pairs(p, q, b, withwt=True), True)
        for c in (D1, R1, D2, R2):
            c.destroy()


def test_ascii_file_read_by_chunks(gpu, tmp_path):
    """The reader's loop in miniature: a text catalogue is read in chunks of bytes, complete lines are parsed, the rest
    is carried to the next chunk (read_ascii.c:750-760, 950-952) and every parsed block goes straight to the device."""
    F = gpu
    x, y, z = box_catalog(60000, 500.0, 17, weights=False)
    path = tmp_path / "cat.txt"
    np.savetxt(path, np.column_stack([x, y, z]), fmt="%.6f")
    b = F.Bins(periodic=True, prec="float", arith=1, **BOX_KW)
    st = F.CatalogStream(bins=b)
    rest = b""
    with open(path, "rb") as fp:
        while True:
            chunk = fp.read(1 << 18)
            if not chunk:
                break
            data = rest + chunk
            cut = data.rfind(b"\n") + 1
            rest = data[cut:]
            if cut:
                rows = np.loadtxt(io.BytesIO(data[:cut]), ndmin=2)
                st.append(rows[:, 0], rows[:, 1], rows[:, 2])
    assert rest == b""
    cat = st.finish()
    one = F.Catalog(x, y, z, bins=b)
    np.testing.assert_array_equal(F.count_pairs(cat, None, b), F.count_pairs(one, None, b))
    cat.destroy()
    one.destroy()


def test_stream_edge_cases_and_errors(gpu):
    F = gpu
    b = F.Bins(periodic=True, prec="double", **BOX_KW)
    # empty catalogue: a valid handle that counts nothing
    st = F.CatalogStream(bins=b)
    st.append([], [], [])
    empty = st.finish()
    assert empty.n == 0
    assert F.count_pairs(empty, None, b).sum() == 0
    x, y, z = box_catalog(2000, 500.0, 3, weights=False)
    one = F.Catalog(x, y, z, bins=b)
    assert F.count_pairs(one, empty, b).sum() == 0
    empty.destroy()
    # a finished stream is gone
    with pytest.raises(F.FcfcGpuError):
        st.append(x, y, z)
    # weights must come with every chunk of a weighted stream, and only then
    sw = F.CatalogStream(bins=b, weighted=True)
    with pytest.raises(ValueError):
        sw.append(x, y, z)
    sw.abort()
    su = F.CatalogStream(bins=b)
    with pytest.raises(ValueError):
        su.append(x, y, z, np.ones_like(x))
    with pytest.raises(ValueError):
        su.append(x, y[:-1], z)
    # non-finite coordinates are reported by finish(), as by the one-shot upload
    xb = x.copy()
    xb[17] = np.nan
    su.append(xb, y, z)
    with pytest.raises(F.FcfcGpuError, match="non-finite"):
        su.finish()
    # the C ABI itself: NULL columns
    L = F.lib()
    h = L.fcfc_gpu_catalog_stream_begin(0, 0, 0)
    assert h
    assert L.fcfc_gpu_catalog_stream_append(h, None, None, None, None, 5) != 0
    assert L.fcfc_gpu_catalog_stream_append(h, None, None, None, None, 0) == 0
    L.fcfc_gpu_catalog_stream_abort(h)
    assert L.fcfc_gpu_catalog_stream_append(None, None, None, None, None, 0) != 0
    one.destroy()
