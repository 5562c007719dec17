"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU restatement (oracle/fcfc_oracle.c).

Only tests/, bench.py's ``cpu_baseline`` leg and ``__graft_entry__.smoke()`` may import this module;
the product path (fcfc_b200/) never does.  See the header of fcfc_oracle.c for what is restated and
how the restatement is pinned against the compiled reference.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = ORACLE_DIR / "liboracle.so"
    srcs = [ORACLE_DIR / "fcfc_oracle.c", ORACLE_DIR / "oracle_impl.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fopenmp", "-shared",
                               "-fPIC", "-o", str(so), str(srcs[0]), "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
    return _LIB


def _bins_struct(real):
    class S(C.Structure):
        _fields_ = [("bintype", C.c_int), ("periodic", C.c_int), ("tabtype", C.c_int),
                    ("ns", C.c_int), ("np", C.c_int), ("nmu", C.c_int),
                    ("swidth", C.c_int), ("pwidth", C.c_int), ("with_mu_one", C.c_int), ("arith", C.c_int),
                    ("rescale", real), ("bsize", real * 3),
                    ("s2bin", C.POINTER(real)), ("pbin", C.POINTER(real)),
                    ("stab", C.c_void_p), ("nstab", C.c_size_t),
                    ("ptab", C.c_void_p), ("nptab", C.c_size_t),
                    ("mutab", C.POINTER(C.c_uint8))]
    return S


_S = {"d": _bins_struct(C.c_double), "f": _bins_struct(C.c_float)}
_NP = {"d": np.float64, "f": np.float32}


def n_linear_bins(lo: float, hi: float, step: float) -> int:
    """Number of bins as counted by the reference (fcfc/2pt_box/load_conf.c:986-995)."""
    n, s = 0, lo
    while s < hi - 1e-10:
        s += step
        n += 1
    return n


@dataclass
class OracleBins:
    prec: str               # 'd' | 'f'
    handle: object          # ctypes pointer owned by the C library
    bintype: int
    periodic: bool
    tabtype: int
    ns: int
    np_: int
    nmu: int
    swidth: int
    pwidth: int
    rescale: float
    bsize: np.ndarray
    s2bin: np.ndarray
    pbin: np.ndarray | None
    stab: np.ndarray
    ptab: np.ndarray | None
    mutab: np.ndarray | None

    @property
    def ntot(self) -> int:
        return self.ns * (1 if self.bintype == 0 else (self.nmu if self.bintype == 1 else self.np_))


def setup(*, prec="d", periodic=True, bintype=0, smin=0.0, smax=200.0, ds=5.0, nmu=1,
          pmin=0.0, pmax=0.0, dpi=0.0, sbin_edges=None, pbin_edges=None, box=None,
          with_mu_one=False, arith=0) -> OracleBins:
    lib = _lib()
    S = _S[prec]
    fn = getattr(lib, f"oracle_setup_{prec}")
    fn.restype = C.POINTER(S)
    lin = sbin_edges is None and (bintype != 2 or pbin_edges is None)
    if sbin_edges is None:
        ns = n_linear_bins(smin, smax, ds)
        sedge = np.array([smin + ds * i for i in range(ns + 1)], dtype=np.float64)
    else:
        sedge = np.ascontiguousarray(sbin_edges, dtype=np.float64)
        ns = len(sedge) - 1
    np_ = 0
    pedge = np.zeros(1)
    if bintype == 2:
        if pbin_edges is None:
            np_ = n_linear_bins(pmin, pmax, dpi)
            pedge = np.array([pmin + dpi * i for i in range(np_ + 1)], dtype=np.float64)
        else:
            pedge = np.ascontiguousarray(pbin_edges, dtype=np.float64)
            np_ = len(pedge) - 1
    b3 = np.ascontiguousarray(np.broadcast_to(np.asarray(box if box is not None else 0.0, dtype=np.float64), (3,)))
    dp = C.POINTER(C.c_double)
    h = fn(C.c_int(int(periodic)), C.c_int(bintype), C.c_int(int(lin)), C.c_double(smin), C.c_double(ds),
           C.c_double(pmin), C.c_double(dpi), sedge.ctypes.data_as(dp), C.c_int(ns),
           pedge.ctypes.data_as(dp), C.c_int(np_), C.c_int(nmu), b3.ctypes.data_as(dp),
           C.c_int(int(with_mu_one)), C.c_int(arith))
    if not h:
        raise RuntimeError("oracle_setup failed")
    s = h.contents
    npd = _NP[prec]

    def arr(p, n):
        return np.ctypeslib.as_array(p, shape=(n,)).astype(npd).copy() if p else None

    def tab(p, n, w):
        if not p:
            return None
        ct = C.c_uint8 if w == 0 else C.c_uint16
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).copy()

    return OracleBins(prec, h, s.bintype, bool(s.periodic), s.tabtype, s.ns, s.np, s.nmu, s.swidth, s.pwidth,
                      float(s.rescale), np.array(list(s.bsize), dtype=npd), arr(s.s2bin, s.ns + 1),
                      arr(s.pbin, s.np + 1) if bintype == 2 else None, tab(s.stab, s.nstab, s.swidth),
                      tab(s.ptab, s.nptab, s.pwidth) if bintype == 2 else None,
                      np.ctypeslib.as_array(s.mutab, shape=(s.nmu * s.nmu,)).copy() if bintype == 1 else None)


def preprocess(bins: OracleBins, cat, arith=None):
    """(x, y, z[, w]) unrescaled -> dict of `real` arrays as tree_create leaves them."""
    lib = _lib()
    npd = _NP[bins.prec]
    x, y, z = (np.array(c, dtype=npd, copy=True) for c in cat[:3])
    w = np.array(cat[3], dtype=npd, copy=True) if len(cat) > 3 else None
    need_s = (not bins.periodic) and bins.bintype != 0
    s = np.zeros_like(x) if need_s else None
    fn = getattr(lib, f"oracle_preprocess_{bins.prec}")
    real = C.c_double if bins.prec == "d" else C.c_float
    rp = C.POINTER(real)
    fn(x.ctypes.data_as(rp), y.ctypes.data_as(rp), z.ctypes.data_as(rp),
       s.ctypes.data_as(rp) if need_s else None, C.c_size_t(len(x)), real(bins.rescale),
       C.c_int(bins.handle.contents.arith if arith is None else arith))
    return {"x": x, "y": y, "z": z, "s": s, "w": w}


def count(bins: OracleBins, cat1, cat2=None, *, withwt=False):
    """Raw count_pairs output for pre-processed catalogues (dicts from preprocess()).
    cat2=None -> auto count (each unordered pair once; the caller doubles)."""
    lib = _lib()
    isauto = cat2 is None
    if isauto:
        cat2 = cat1
    real = C.c_double if bins.prec == "d" else C.c_float
    rp = C.POINTER(real)

    def p(a):
        return a.ctypes.data_as(rp) if a is not None else None

    ci = np.zeros(bins.ntot, dtype=np.int64)
    cd = np.zeros(bins.ntot, dtype=np.float64)
    w1, w2 = cat1["w"], cat2["w"]
    if withwt:
        if w1 is None:
            w1 = np.ones_like(cat1["x"])
        if w2 is None:
            w2 = np.ones_like(cat2["x"])
    fn = getattr(lib, f"oracle_count_{bins.prec}")
    fn.restype = C.c_int
    rc = fn(bins.handle, p(cat1["x"]), p(cat1["y"]), p(cat1["z"]), p(cat1["s"]), p(w1), C.c_size_t(len(cat1["x"])),
            p(cat2["x"]), p(cat2["y"]), p(cat2["z"]), p(cat2["s"]), p(w2), C.c_size_t(len(cat2["x"])),
            C.c_int(int(isauto)), C.c_int(int(withwt)),
            ci.ctypes.data_as(C.POINTER(C.c_int64)), cd.ctypes.data_as(C.POINTER(C.c_double)))
    if rc:
        raise RuntimeError("oracle_count failed")
    return cd if withwt else ci
