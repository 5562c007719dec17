"""CPU: the C restatement (oracle/) against the golden vectors produced by the compiled reference.

This is what pins the oracle (the reference ships no tests of its own)."""
import hashlib

import numpy as np
import pytest

from cases import CASES, make_catalog
from conftest import load_golden
from oracle import oracle


def _sha(cats):
    h = hashlib.sha256()
    for c in cats:
        for a in c:
            h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("prec", ["dbl", "flt"])
def test_oracle_matches_reference_golden(name, prec):
    case = CASES[name]
    g = load_golden(name)
    cats = [make_catalog(s, case["withwt"]) for s in case["cats"]]
    assert _sha(cats) == str(g["input_sha256"]), "catalogue generator drifted from the golden inputs"
    ob = oracle.setup(prec=prec[0], periodic=case["periodic"], **case["kw"])
    # tables and rescale factor exactly as cf_setup built them
    assert ob.rescale == float(g[f"{prec}_rescale"])
    assert ob.tabtype == int(g[f"{prec}_tabtype"])
    np.testing.assert_array_equal(ob.s2bin.astype(np.float64), g[f"{prec}_s2bin"])
    np.testing.assert_array_equal(ob.stab, g[f"{prec}_stab"])
    if ob.bintype == 2:
        np.testing.assert_array_equal(ob.pbin.astype(np.float64), g[f"{prec}_pbin"])
        np.testing.assert_array_equal(ob.ptab, g[f"{prec}_ptab"])
    if ob.bintype == 1:
        np.testing.assert_array_equal(ob.mutab, g[f"{prec}_mutab"])
    if case["periodic"]:
        np.testing.assert_array_equal(ob.bsize.astype(np.float64), g[f"{prec}_bsize"])
    pc = [oracle.preprocess(ob, c) for c in cats]
    for p in case["pairs"]:
        i, j = "DR".index(p[0]), "DR".index(p[1])
        c = oracle.count(ob, pc[i], None if i == j else pc[j], withwt=case["withwt"])
        ref = g[f"{prec}_scalar_{p}"]
        if case["withwt"]:
            # weighted sums: <= 1e-12 relative (summation order differs), identical empty-bin pattern
            np.testing.assert_array_equal(c == 0, ref == 0)
            np.testing.assert_allclose(c, ref, rtol=1e-12, atol=0)
        else:
            np.testing.assert_array_equal(c, ref)       # bit-exact


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if not CASES[n]["withwt"]])
def test_reference_simd_equals_scalar_in_double(name):
    """Recorded fact: in double precision the AVX-512 build gives the same counts."""
    g = load_golden(name)
    for p in CASES[name]["pairs"]:
        np.testing.assert_array_equal(g[f"dbl_avx512_{p}"], g[f"dbl_scalar_{p}"])


def test_lattice_float_equals_double():
    """On the 2^-3 lattice with power-of-two bins all arithmetic is exact: float and double agree."""
    g = load_golden("box_lattice_smu")
    np.testing.assert_array_equal(g["flt_scalar_DD"], g["dbl_scalar_DD"])


def test_fma_mode_equals_scalar_on_lattice():
    case = CASES["box_lattice_smu"]
    cats = [make_catalog(s, False) for s in case["cats"]]
    for prec in "df":
        out = []
        for arith in (0, 1):
            ob = oracle.setup(prec=prec, periodic=True, arith=arith, **case["kw"])
            out.append(oracle.count(ob, oracle.preprocess(ob, cats[0])))
        np.testing.assert_array_equal(out[0], out[1])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_fma_order_matches_reference_avx512_double(name):
    """The FMA evaluation order (arith = 1: fused distance chain, (num * nmu^2) / d^2 rounded toward zero,
    metric_common.c:377-534) is the mode bench.py times.  Its pin is the reference's AVX-512 build in double
    precision: exact for unweighted counts, 1e-12 for weighted sums."""
    case = CASES[name]
    g = load_golden(name)
    cats = [make_catalog(s, case["withwt"]) for s in case["cats"]]
    ob = oracle.setup(prec="d", periodic=case["periodic"], arith=1, **case["kw"])
    pc = [oracle.preprocess(ob, c) for c in cats]
    for p in case["pairs"]:
        i, j = "DR".index(p[0]), "DR".index(p[1])
        c = oracle.count(ob, pc[i], None if i == j else pc[j], withwt=case["withwt"])
        ref = g[f"dbl_avx512_{p}"]
        if case["withwt"]:
            np.testing.assert_array_equal(c == 0, ref == 0)
            np.testing.assert_allclose(c, ref, rtol=1e-12, atol=0)
        else:
            np.testing.assert_array_equal(c, ref)


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if not CASES[n]["withwt"]])
def test_oracle_fma_order_within_reference_float_spread(name):
    """SINGLE_PREC + AVX-512 is not one formula in the reference: vector blocks use the FMA chain, odd remainders the
    scalar code (metric_common.c:1436-1473), periodic pairs either the node shift or the per-pair wrap.  The FMA-order
    restatement is therefore bounded, not pinned, in float: per bin it must differ from the reference's AVX-512 float
    build by no more than that build differs from the reference's own scalar float build."""
    case = CASES[name]
    g = load_golden(name)
    cats = [make_catalog(s, False) for s in case["cats"]]
    ob = oracle.setup(prec="f", periodic=case["periodic"], arith=1, **case["kw"])
    pc = [oracle.preprocess(ob, c) for c in cats]
    for p in case["pairs"]:
        i, j = "DR".index(p[0]), "DR".index(p[1])
        c = oracle.count(ob, pc[i], None if i == j else pc[j])
        simd, scal = g[f"flt_avx512_{p}"], g[f"flt_scalar_{p}"]
        own, ours = np.abs(simd - scal), np.abs(c - simd)
        assert ours.sum() <= max(8, 2 * own.sum()) and ours.max() <= max(2, 2 * own.max()), (ours.sum(), own.sum())
        assert abs(int(c.sum()) - int(simd.sum())) <= max(2, 2 * abs(int(simd.sum()) - int(scal.sum())))
