"""Python host mirror of the reference's counting interface, on top of the C ABI (include/fcfc_gpu.h).

The names follow the reference's seam:

  =====================================  ==============================================================
  reference (src/fcfc/2pt_box/...)        here
  =====================================  ==============================================================
  cf_setup()        setup_cf.c:547        :class:`Bins` (rescale factor, rescaled edges, lookup tables)
  tree_create()     build_tree.c:36       :class:`Catalog` (rescale + upload + cell list on the GPU)
  count_pairs()     count_func.c:4847     :func:`count_pairs` (raw counts, auto pairs counted once)
  tree_destroy()    build_tree.c:230      :meth:`Catalog.destroy`
  =====================================  ==============================================================

Everything numerical happens inside ``libfcfc_b200.so`` on the GPU.  There is no CPU fallback: if the
library is missing or no sm_100 device is usable, calls raise :class:`FcfcGpuError`.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("FCFC_B200_LIB", PKG / "libfcfc_b200.so"))     # (the override serves A/B builds: tools/build_variant.py)

BIN_ISO, BIN_SMU, BIN_SPI = 0, 1, 2
ARITH_SCALAR, ARITH_FMA = 0, 1


class FcfcGpuError(RuntimeError):
    def __init__(self, msg, code=None):
        super().__init__(msg)
        self.code = code


class _CBins(C.Structure):
    _fields_ = [("bintype", C.c_int32), ("periodic", C.c_int32), ("is_float", C.c_int32), ("tabtype", C.c_int32),
                ("ns", C.c_int32), ("np", C.c_int32), ("nmu", C.c_int32), ("swidth", C.c_int32),
                ("pwidth", C.c_int32), ("with_mu_one", C.c_int32), ("arith", C.c_int32), ("reserved", C.c_int32),
                ("s2bin", C.c_void_p), ("pbin", C.c_void_p), ("stab", C.c_void_p), ("ptab", C.c_void_p),
                ("mutab", C.c_void_p), ("nstab", C.c_uint64), ("nptab", C.c_uint64), ("bsize", C.c_double * 3)]


class _CStats(C.Structure):
    _fields_ = [("pair_evals", C.c_uint64), ("pairs_in", C.c_uint64), ("ms_sort", C.c_double),
                ("ms_count", C.c_double), ("ms_total", C.c_double), ("kernel_launches", C.c_uint32),
                ("ncell", C.c_int32 * 3), ("nitem", C.c_int32), ("dense_rows", C.c_int32), ("prefilter", C.c_int32),
                ("classified", C.c_int32), ("pair_evals_computed", C.c_uint64)]


_lib = None


def lib():
    """Load the CUDA library; never falls back to anything else."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FcfcGpuError(f"{LIB_PATH} is missing: build it with `python -m fcfc_b200.build` "
                               "(fcfc_b200 has no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        L.fcfc_gpu_last_error.restype = C.c_char_p
        L.fcfc_gpu_catalog_create.restype = C.c_void_p
        L.fcfc_gpu_catalog_create.argtypes = [C.c_void_p] * 5 + [C.c_size_t, C.c_int, C.c_double, C.c_int]
        L.fcfc_gpu_catalog_destroy.argtypes = [C.c_void_p]
        L.fcfc_gpu_catalog_size.restype = C.c_size_t
        L.fcfc_gpu_catalog_size.argtypes = [C.c_void_p]
        L.fcfc_gpu_catalog_wsum.restype = C.c_double
        L.fcfc_gpu_catalog_wsum.argtypes = [C.c_void_p]
        L.fcfc_gpu_catalog_stream_begin.restype = C.c_void_p
        L.fcfc_gpu_catalog_stream_begin.argtypes = [C.c_size_t, C.c_int, C.c_int]
        L.fcfc_gpu_catalog_stream_append.argtypes = [C.c_void_p] * 5 + [C.c_size_t]
        L.fcfc_gpu_catalog_stream_size.restype = C.c_size_t
        L.fcfc_gpu_catalog_stream_size.argtypes = [C.c_void_p]
        L.fcfc_gpu_catalog_stream_finish.restype = C.c_void_p
        L.fcfc_gpu_catalog_stream_finish.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.fcfc_gpu_catalog_stream_abort.argtypes = [C.c_void_p]
        L.fcfc_gpu_count_partial.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_CBins), C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fcfc_gpu_count.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_CBins), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fcfc_gpu_bins_create.restype = C.c_void_p
        L.fcfc_gpu_bins_create.argtypes = [C.c_int] * 4 + [C.c_double] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                                               C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.fcfc_gpu_bins_get.restype = C.POINTER(_CBins)
        L.fcfc_gpu_bins_get.argtypes = [C.c_void_p]
        L.fcfc_gpu_bins_rescale.restype = C.c_double
        L.fcfc_gpu_bins_rescale.argtypes = [C.c_void_p]
        L.fcfc_gpu_bins_free.argtypes = [C.c_void_p]
        L.fcfc_gpu_measure_fp32_peak.restype = C.c_double
        L.fcfc_gpu_measure_fp32_peak.argtypes = [C.POINTER(C.c_double)]
        L.fcfc_gpu_measure_fp64_peak.restype = C.c_double
        L.fcfc_gpu_measure_fp64_peak.argtypes = []
        L.fcfc_gpu_measure_smem_atomic_peak.restype = C.c_double
        L.fcfc_gpu_measure_smem_atomic_peak.argtypes = []
        L.fcfc_gpu_init.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.fcfc_gpu_set_option.argtypes = [C.c_char_p, C.c_long]
        _lib = L
    return _lib


def last_error() -> str:
    return lib().fcfc_gpu_last_error().decode(errors="replace")


def init(devices=None, verbose: int = 0) -> int:
    """Bind to CUDA devices (default: all visible).  Raises when none is usable."""
    L = lib()
    if devices is None:
        n = L.fcfc_gpu_init(0, None, verbose)
    else:
        arr = (C.c_int * len(devices))(*devices)
        n = L.fcfc_gpu_init(len(devices), arr, verbose)
    if n <= 0:
        raise FcfcGpuError(last_error(), n)
    return n


def set_option(name: str, value: int = 1) -> None:
    """Tuning / diagnostic switch of the engine (include/fcfc_gpu.h: fcfc_gpu_set_option); "defaults" resets all."""
    if lib().fcfc_gpu_set_option(name.encode(), int(value)) != 0:
        raise FcfcGpuError(last_error(), -2)


def n_linear_bins(lo: float, hi: float, step: float) -> int:
    """Bin count as the reference derives it from MIN/MAX/SIZE (load_conf.c:986-995)."""
    n, s = 0, lo
    while s < hi - 1e-10:
        s += step
        n += 1
    return n


class Bins:
    """The slice of the reference's ``CF`` that count_pairs reads (eval_cf.h:54-67), built as cf_setup does."""

    def __init__(self, *, periodic: bool, prec: str = "double", bintype: int = BIN_ISO,
                 smin: float = 0.0, smax: float | None = None, ds: float | None = None, nmu: int = 1,
                 pmin: float = 0.0, pmax: float | None = None, dpi: float | None = None,
                 sbin_edges=None, pbin_edges=None, box=None, with_mu_one: bool = False, arith: int = ARITH_SCALAR):
        L = lib()
        self.is_float = prec in ("float", "f", "flt", "single")
        self.dtype = np.float32 if self.is_float else np.float64
        linear = sbin_edges is None and (bintype != BIN_SPI or pbin_edges is None)
        if sbin_edges is None:
            ns = n_linear_bins(smin, smax, ds)
            sedge = None
        else:
            sedge = np.ascontiguousarray(sbin_edges, dtype=np.float64)
            ns = len(sedge) - 1
            if not linear and bintype == BIN_SPI and pbin_edges is None:
                pbin_edges = [pmin + dpi * i for i in range(n_linear_bins(pmin, pmax, dpi) + 1)]
        np_, pedge = 0, None
        if bintype == BIN_SPI:
            if pbin_edges is None:
                np_ = n_linear_bins(pmin, pmax, dpi)
            else:
                pedge = np.ascontiguousarray(pbin_edges, dtype=np.float64)
                np_ = len(pedge) - 1
                if sedge is None:
                    sedge = np.array([smin + ds * i for i in range(ns + 1)], dtype=np.float64)
        if periodic:
            if box is None:
                raise ValueError("periodic counts need BOX_SIZE")
            b3 = np.ascontiguousarray(np.broadcast_to(np.asarray(box, dtype=np.float64), (3,)))
        else:
            b3 = np.zeros(3)
        self._keep = (sedge, pedge, b3)
        self._h = L.fcfc_gpu_bins_create(int(periodic), int(self.is_float), bintype, int(linear),
                                         float(smin), float(ds or 0), float(pmin), float(dpi or 0),
                                         sedge.ctypes.data if sedge is not None else None, ns,
                                         pedge.ctypes.data if pedge is not None else None, np_,
                                         int(nmu), b3.ctypes.data, int(with_mu_one), int(arith))
        if not self._h:
            raise FcfcGpuError("fcfc_gpu_bins_create failed: invalid binning")
        self.c = L.fcfc_gpu_bins_get(self._h)
        self.rescale = L.fcfc_gpu_bins_rescale(self._h)
        c = self.c.contents
        self.periodic, self.bintype, self.tabtype = bool(c.periodic), c.bintype, c.tabtype
        self.ns, self.np_, self.nmu = c.ns, c.np, c.nmu
        self.arith = c.arith

    @property
    def ntot(self) -> int:
        return self.ns * (1 if self.bintype == BIN_ISO else (self.nmu if self.bintype == BIN_SMU else self.np_))

    # read-only views for tests / inspection
    def _arr(self, ptr, n, dt):
        if not ptr or n == 0:
            return None
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()

    @property
    def s2bin(self):
        return self._arr(self.c.contents.s2bin, self.ns + 1, self.dtype)

    @property
    def pbin(self):
        return self._arr(self.c.contents.pbin, self.np_ + 1, self.dtype) if self.bintype == BIN_SPI else None

    @property
    def stab(self):
        c = self.c.contents
        return self._arr(c.stab, c.nstab, np.uint16 if c.swidth else np.uint8)

    @property
    def ptab(self):
        c = self.c.contents
        return self._arr(c.ptab, c.nptab, np.uint16 if c.pwidth else np.uint8) if self.bintype == BIN_SPI else None

    @property
    def mutab(self):
        return self._arr(self.c.contents.mutab, self.nmu * self.nmu, np.uint8) if self.bintype == BIN_SMU else None

    @property
    def bsize(self):
        return np.array(list(self.c.contents.bsize))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().fcfc_gpu_bins_free(self._h)
                self._h = None
        except Exception:
            pass


class Catalog:
    """A catalogue resident on the GPU (replaces the k-d / ball tree handle of tree_create)."""

    def __init__(self, x, y, z, w=None, *, bins: Bins, x2sum=None, prescaled: bool = False):
        """x, y, z[, w]: host arrays of *unrescaled* coordinates (as read from the catalogue file); they are
        converted to the build's ``real`` type and multiplied by ``bins.rescale`` on the device, exactly as
        tree_create does on the host (build_tree.c:121-131).  prescaled=True skips the multiplication."""
        L = lib()
        dt = bins.dtype
        xs = [np.ascontiguousarray(a, dtype=dt) for a in (x, y, z)]
        n = len(xs[0])
        if any(len(a) != n for a in xs):
            raise ValueError("coordinate arrays differ in length")
        wv = np.ascontiguousarray(w, dtype=dt) if w is not None else None
        sv = np.ascontiguousarray(x2sum, dtype=dt) if x2sum is not None else None
        need_s = (not bins.periodic) and bins.bintype != BIN_ISO
        sumsq = -1 if (sv is not None or not need_s) else bins.arith
        self.n, self.has_w, self.is_float = n, wv is not None, bins.is_float
        self._h = L.fcfc_gpu_catalog_create(xs[0].ctypes.data, xs[1].ctypes.data, xs[2].ctypes.data,
                                            sv.ctypes.data if sv is not None else None,
                                            wv.ctypes.data if wv is not None else None, n, int(bins.is_float),
                                            1.0 if prescaled else float(bins.rescale), sumsq)
        if not self._h:
            raise FcfcGpuError(last_error())

    @classmethod
    def from_device(cls, px: int, py: int, pz: int, n: int, *, bins: Bins, pw: int | None = None, px2sum: int | None = None,
                    prescaled: bool = False) -> "Catalog":
        """The same from DEVICE pointers to columns of the build's ``real`` type (e.g. slices all-gathered over NVLink
        by a one-process-per-GPU launcher, fcfc_b200/sharding.py): the C ABI resolves host and device pointers alike."""
        self = cls.__new__(cls)
        need_s = (not bins.periodic) and bins.bintype != BIN_ISO
        sumsq = -1 if (px2sum is not None or not need_s) else bins.arith
        self.n, self.has_w, self.is_float = n, pw is not None, bins.is_float
        self._h = lib().fcfc_gpu_catalog_create(px, py, pz, px2sum, pw, n, int(bins.is_float),
                                                1.0 if prescaled else float(bins.rescale), sumsq)
        if not self._h:
            raise FcfcGpuError(last_error())
        return self

    @property
    def wsum(self) -> float:
        return lib().fcfc_gpu_catalog_wsum(self._h)

    def destroy(self):
        if getattr(self, "_h", None):
            lib().fcfc_gpu_catalog_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class CatalogStream:
    """A catalogue built chunk by chunk while the host is still reading the file (include/fcfc_gpu.h:
    fcfc_gpu_catalog_stream_*): every :meth:`append` copies its rows to pinned staging memory and returns, the transfer
    to the device overlaps the parsing of the next chunk (the chunk loop of io/read_ascii.c:750-950), and
    :meth:`finish` returns the same :class:`Catalog` the one-shot constructor builds from the concatenated columns."""

    def __init__(self, *, bins: Bins, weighted: bool = False, n_hint: int = 0, prescaled: bool = False):
        self._bins, self._weighted, self._prescaled = bins, bool(weighted), prescaled
        self._h = lib().fcfc_gpu_catalog_stream_begin(int(n_hint), int(bins.is_float), int(weighted))
        if not self._h:
            raise FcfcGpuError(last_error())

    def __len__(self) -> int:
        return lib().fcfc_gpu_catalog_stream_size(self._h) if self._h else 0

    def append(self, x, y, z, w=None) -> None:
        if not self._h:
            raise FcfcGpuError("catalogue stream already finished")
        dt = self._bins.dtype
        xs = [np.ascontiguousarray(a, dtype=dt) for a in (x, y, z)]
        n = len(xs[0])
        if any(len(a) != n for a in xs):
            raise ValueError("coordinate arrays differ in length")
        if self._weighted != (w is not None):
            raise ValueError("every chunk of a weighted stream needs weights (and only those)")
        wv = np.ascontiguousarray(w, dtype=dt) if w is not None else None
        if wv is not None and len(wv) != n:
            raise ValueError("weights differ in length from the coordinates")
        rc = lib().fcfc_gpu_catalog_stream_append(self._h, xs[0].ctypes.data, xs[1].ctypes.data, xs[2].ctypes.data,
                                                  wv.ctypes.data if wv is not None else None, n)
        if rc != 0:
            raise FcfcGpuError(last_error(), rc)

    def finish(self) -> Catalog:
        if not self._h:
            raise FcfcGpuError("catalogue stream already finished")
        b = self._bins
        n = len(self)
        need_s = (not b.periodic) and b.bintype != BIN_ISO
        h, self._h = self._h, None
        cat = Catalog.__new__(Catalog)
        cat.n, cat.has_w, cat.is_float = n, self._weighted, b.is_float
        cat._h = lib().fcfc_gpu_catalog_stream_finish(h, 1.0 if self._prescaled else float(b.rescale), b.arith if need_s else -1)
        if not cat._h:
            raise FcfcGpuError(last_error())
        return cat

    def abort(self) -> None:
        if getattr(self, "_h", None):
            lib().fcfc_gpu_catalog_stream_abort(self._h)
            self._h = None

    def __del__(self):
        try:
            self.abort()
        except Exception:
            pass


def count_pairs(cat1: Catalog, cat2: Catalog | None, bins: Bins, *, withwt: bool = False,
                part: int = 0, nparts: int = 1, dev_hist_ptr: int | None = None) -> np.ndarray:
    """count_pairs(): raw pair counts (int64) or weighted sums (float64), ``ntot`` entries laid out as
    cnt[s + p*ns].  cat2=None (or cat2 is cat1) -> auto count: every unordered pair once, the caller doubles
    (eval_cf.c:126-133).  part/nparts select one shard of the primary catalogue's work items."""
    L = lib()
    isauto = cat2 is None or cat2 is cat1
    c2 = cat1 if isauto else cat2
    out = np.zeros(bins.ntot, dtype=np.float64 if withwt else np.int64)
    if nparts == 1 and dev_hist_ptr is None:
        # the whole count, over every device the process is bound to (one all-reduce of the histograms)
        rc = L.fcfc_gpu_count(cat1._h, c2._h, bins.c, int(isauto), int(withwt),
                              None if withwt else out.ctypes.data, out.ctypes.data if withwt else None)
    else:
        rc = L.fcfc_gpu_count_partial(cat1._h, c2._h, bins.c, int(isauto), int(withwt), part, nparts,
                                      None if withwt else out.ctypes.data, out.ctypes.data if withwt else None,
                                      dev_hist_ptr)
    if rc != 0:
        raise FcfcGpuError(last_error(), rc)
    return out


def stats() -> dict:
    s = _CStats()
    lib().fcfc_gpu_get_stats(C.byref(s))
    return {"pair_evals": s.pair_evals, "pairs_in": s.pairs_in, "ms_sort": s.ms_sort, "ms_count": s.ms_count,
            "ms_total": s.ms_total, "kernel_launches": s.kernel_launches, "ncell": list(s.ncell), "nitem": s.nitem,
            "dense_rows": s.dense_rows, "prefilter": s.prefilter, "classified": s.classified,
            "pair_evals_computed": s.pair_evals_computed}


def measure_fp32_peak() -> tuple[float, float]:
    """(FP32 lane-instructions per second, implied SM clock in MHz at 128 lanes/SM)."""
    clk = C.c_double(0)
    v = lib().fcfc_gpu_measure_fp32_peak(C.byref(clk))
    if v <= 0:
        raise FcfcGpuError("FP32 peak measurement failed (no device?)")
    return v, clk.value


def measure_fp64_peak() -> float:
    """FP64 lane-instructions per second (DFMA stream on all SMs)."""
    v = lib().fcfc_gpu_measure_fp64_peak()
    if v <= 0:
        raise FcfcGpuError("FP64 peak measurement failed (no device?)")
    return v


def measure_smem_atomic_peak() -> float:
    """Shared-memory histogram increments per second (red.shared.add.u32 lanes, whole device)."""
    v = lib().fcfc_gpu_measure_smem_atomic_peak()
    if v <= 0:
        raise FcfcGpuError("shared-memory atomic peak measurement failed (no device?)")
    return v
