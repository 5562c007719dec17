"""CPU: the product's host helper fcfc_gpu_bins_create (cf_setup mirror) against the reference's tables."""
import numpy as np
import pytest

import fcfc_b200 as F
from cases import CASES
from conftest import load_golden
from oracle import oracle


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("prec", ["dbl", "flt"])
def test_bins_match_reference_golden(name, prec):
    case = CASES[name]
    g = load_golden(name)
    b = F.Bins(periodic=case["periodic"], prec="float" if prec == "flt" else "double", **case["kw"])
    assert b.rescale == float(g[f"{prec}_rescale"])
    assert b.tabtype == int(g[f"{prec}_tabtype"])
    np.testing.assert_array_equal(b.s2bin.astype(np.float64), g[f"{prec}_s2bin"])
    np.testing.assert_array_equal(b.stab, g[f"{prec}_stab"])
    if b.bintype == F.BIN_SPI:
        np.testing.assert_array_equal(b.pbin.astype(np.float64), g[f"{prec}_pbin"])
        np.testing.assert_array_equal(b.ptab, g[f"{prec}_ptab"])
    if b.bintype == F.BIN_SMU:
        np.testing.assert_array_equal(b.mutab, g[f"{prec}_mutab"])
    if case["periodic"]:
        np.testing.assert_array_equal(b.bsize, g[f"{prec}_bsize"])


@pytest.mark.parametrize("kw", [
    dict(bintype=0, smin=0.5, smax=150.5, ds=0.3),              # 1 decimal digit -> factor 10/3
    dict(bintype=0, smin=0.0, smax=200.0, ds=0.01),             # integer table too long -> hybrid
    dict(bintype=0, smin=0.0, smax=30.0, ds=0.125),             # power-of-two step
    dict(bintype=2, smin=1.0, smax=41.0, ds=2.5, pmin=0.0, pmax=60.0, dpi=1.5),
    dict(bintype=1, smin=0.0, smax=150.0, ds=1.0, nmu=255),
    dict(bintype=0, sbin_edges=[0.1, 0.5, 2.0, 9.0, 33.3, 120.0]),
])
@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("prec", ["double", "float"])
def test_bins_match_oracle(kw, periodic, prec):
    box = dict(box=[1000.0, 1200.0, 900.0]) if periodic else {}
    b = F.Bins(periodic=periodic, prec=prec, **kw, **box)
    o = oracle.setup(prec=prec[0], periodic=periodic, **kw, **box)
    assert b.rescale == o.rescale and b.tabtype == o.tabtype and b.ns == o.ns and b.np_ == o.np_
    np.testing.assert_array_equal(b.s2bin, o.s2bin)
    np.testing.assert_array_equal(b.stab, o.stab)
    if b.bintype == 2:
        np.testing.assert_array_equal(b.pbin, o.pbin)
        np.testing.assert_array_equal(b.ptab, o.ptab)
    if b.bintype == 1:
        np.testing.assert_array_equal(b.mutab, o.mutab)
        np.testing.assert_array_equal(b.mutab, np.floor(np.sqrt(np.arange(b.nmu ** 2))).astype(np.uint8))
    if periodic:
        np.testing.assert_array_equal(b.bsize.astype(b.dtype), o.bsize)


def test_bin_count_rule():
    assert F.n_linear_bins(0, 200, 5) == 40
    assert F.n_linear_bins(0, 0.3, 0.1) == 3       # accumulated 0.1+0.1+0.1 < 0.3 - tol is false
    assert F.n_linear_bins(10, 150, 2.5) == 56
