"""TEST INFRASTRUCTURE ONLY -- runs the compiled, unmodified FCFC reference (oracle/_ref).

Only tests/, bench.py's CPU arms (``--impl reference`` / ``cpu_baseline``) and
``__graft_entry__.smoke()`` may import this module.  The product (fcfc_b200/) never does.

It drives ``oracle/_ref/<flavour>/ref_driver_{box,svy}`` (oracle/ref_driver.c linked against the
reference's own objects): writes catalogues in the driver's ``.fbin`` container, writes an FCFC
configuration file with the reference's own keywords (reference etc/fcfc_2pt_box.conf,
etc/fcfc_2pt.conf), runs the binary and parses the ``.drv`` dump (raw ``count_pairs`` output, its
wall time, and the bin tables ``cf_setup`` built).
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REF_DIR = ORACLE_DIR / "_ref"


def cpu_flags() -> set[str]:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def best_simd_flavour(prec: str) -> str:
    """Fastest reference flavour the host CPU can execute ('dbl'|'flt')."""
    fl = cpu_flags()
    if {"avx512f", "avx512dq", "avx512vl", "avx512bw"} <= fl:
        return f"{prec}_avx512"
    if {"avx2", "fma"} <= fl:
        return f"{prec}_avx2"
    return f"{prec}_scalar"


def have_flavour(flavour: str, prog: str = "box") -> bool:
    return (REF_DIR / flavour / f"ref_driver_{prog}").exists()


def write_fbin(path, cols) -> None:
    """cols: sequence of 1-D arrays (x, y, z[, w]); stored as float64 columns."""
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
    n = len(cols[0])
    with open(path, "wb") as f:
        f.write(b"FCFCBIN1")
        f.write(struct.pack("<QII", n, len(cols), 0))
        for c in cols:
            assert len(c) == n
            c.tofile(f)


@dataclass
class PairResult:
    label: str
    isauto: bool
    withwt: bool
    t_tree: float
    t_count: float
    n1: int
    n2: int
    wt1: float
    wt2: float
    cnt: np.ndarray  # int64 (unweighted) or float64 (weighted), raw count_pairs output


@dataclass
class DrvResult:
    is_float: bool
    periodic: bool
    bintype: int
    tabtype: int
    ns: int
    np_: int
    nmu: int
    swidth: int
    pwidth: int
    npc: int
    ncat: int
    nthread: int
    ntot: int
    rescale: float
    bsize: np.ndarray
    s2bin: np.ndarray
    pbin: np.ndarray | None
    stab: np.ndarray
    ptab: np.ndarray | None
    mutab: np.ndarray | None
    pairs: list[PairResult] = field(default_factory=list)
    stdout: str = ""


def parse_drv(path) -> DrvResult:
    buf = Path(path).read_bytes()
    assert buf[:8] == b"FCFCDRV1", "bad driver dump"
    off = 8
    ints = struct.unpack_from("<12i", buf, off); off += 48
    (is_float, periodic, bintype, tabtype, ns, np_, nmu, swidth, pwidth, npc, ncat, nthread) = ints
    (ntot,) = struct.unpack_from("<Q", buf, off); off += 8
    (rescale,) = struct.unpack_from("<d", buf, off); off += 8
    bsize = np.frombuffer(buf, "<f8", 3, off).copy(); off += 24

    def reals():
        nonlocal off
        (n,) = struct.unpack_from("<Q", buf, off); off += 8
        a = np.frombuffer(buf, "<f8", n, off).copy(); off += 8 * n
        return a if n else None

    def table(width):
        nonlocal off
        (n,) = struct.unpack_from("<Q", buf, off); off += 8
        dt = np.uint8 if width == 0 else np.dtype("<u2")
        a = np.frombuffer(buf, dt, n, off).copy(); off += n * np.dtype(dt).itemsize
        return a if n else None

    s2bin = reals()
    pbin = reals()
    stab = table(swidth)
    ptab = table(pwidth)
    mutab = table(0)
    res = DrvResult(bool(is_float), bool(periodic), bintype, tabtype, ns, np_, nmu, swidth, pwidth,
                    npc, ncat, nthread, ntot, rescale, bsize, s2bin, pbin, stab, ptab, mutab)
    for _ in range(npc):
        lab = buf[off:off + 2].decode(); off += 2
        isauto, withwt = struct.unpack_from("<2i", buf, off); off += 8
        t_tree, t_count = struct.unpack_from("<2d", buf, off); off += 16
        n1, n2 = struct.unpack_from("<2Q", buf, off); off += 16
        wt1, wt2 = struct.unpack_from("<2d", buf, off); off += 16
        dt = "<f8" if withwt else "<i8"
        cnt = np.frombuffer(buf, dt, ntot, off).copy(); off += 8 * ntot
        res.pairs.append(PairResult(lab, bool(isauto), bool(withwt), t_tree, t_count, n1, n2, wt1, wt2, cnt))
    return res


def run_reference(catalogs, *, periodic: bool, prec: str = "dbl", isa: str = "scalar",
                  pairs=("DD",), labels=None, box=None, bintype: int = 0,
                  smin: float = 0.0, smax: float = 200.0, ds: float = 5.0, nmu: int = 1,
                  pmin: float = 0.0, pmax: float = 0.0, dpi: float = 0.0,
                  sbin_edges=None, pbin_edges=None, data_struct: int = 0,
                  threads: int | None = None, workdir=None, keep: bool = False,
                  timeout: float = 3600.0) -> DrvResult:
    """Run the unmodified reference on in-memory catalogues.

    catalogs: list of tuples (x, y, z) or (x, y, z, w) of *unrescaled* comoving coordinates.
    pairs: FCFC pair labels, e.g. ("DD", "DR", "RR"); catalogue i is labelled labels[i] (default
    'D','R','S',...).  Binning keywords as in the reference's configuration file.
    """
    flavour = f"{prec}_{isa}"
    prog = "box" if periodic else "svy"
    exe = REF_DIR / flavour / f"ref_driver_{prog}"
    if not exe.exists():
        raise FileNotFoundError(f"{exe} missing: run `make -C oracle` where /root/reference exists")
    labels = labels or "DRSTUVWXYZ"[: len(catalogs)]
    tmp = Path(workdir) if workdir else Path(tempfile.mkdtemp(prefix="fcfc_ref_"))
    tmp.mkdir(parents=True, exist_ok=True)
    files, has_wt = [], []
    for i, cat in enumerate(catalogs):
        p = tmp / f"cat_{labels[i]}.fbin"
        write_fbin(p, cat)
        files.append(str(p))
        has_wt.append(len(cat) == 4)
    lines = []
    lines.append("CATALOG = [" + ",".join(f'"{f}"' for f in files) + "]")
    lines.append("CATALOG_LABEL = [" + ",".join(labels[: len(catalogs)]) + "]")
    fm = "%f" if prec == "flt" else "%lf"
    lines.append("ASCII_FORMATTER = [" + ",".join(f'"{fm} {fm} {fm} {fm}"' for _ in catalogs) + "]")
    lines.append("POSITION = [" + ",".join('"$1","$2","$3"' for _ in catalogs) + "]")
    if any(has_wt):
        lines.append("WEIGHT = [" + ",".join('"$4"' if w else '"1"' for w in has_wt) + "]")
        if not all(has_wt):
            raise ValueError("mixed weighted/unweighted catalogues: give explicit unit weights")
    if periodic:
        b = np.atleast_1d(np.asarray(box, dtype=float))
        lines.append("BOX_SIZE = " + (repr(float(b[0])) if b.size == 1 else "[" + ",".join(repr(float(v)) for v in b) + "]"))
    else:
        lines.append("COORD_CONVERT = [" + ",".join("F" for _ in catalogs) + "]")
    lines.append(f"DATA_STRUCT = {data_struct}")
    lines.append(f"BINNING_SCHEME = {bintype}")
    lines.append("PAIR_COUNT = [" + ",".join(pairs) + "]")
    lines.append("PAIR_COUNT_FILE = [" + ",".join(f'"{tmp}/pc_{p}.bin"' for p in pairs) + "]")
    if sbin_edges is not None:
        e = np.asarray(sbin_edges, dtype=float)
        np.savetxt(tmp / "sbins.txt", np.c_[e[:-1], e[1:]], fmt="%.17g")
        lines.append(f'SEP_BIN_FILE = "{tmp}/sbins.txt"')
    else:
        lines += [f"SEP_BIN_MIN = {smin!r}", f"SEP_BIN_MAX = {smax!r}", f"SEP_BIN_SIZE = {ds!r}"]
    if bintype == 1:
        lines.append(f"MU_BIN_NUM = {nmu}")
    if bintype == 2:
        if pbin_edges is not None:
            e = np.asarray(pbin_edges, dtype=float)
            np.savetxt(tmp / "pbins.txt", np.c_[e[:-1], e[1:]], fmt="%.17g")
            lines.append(f'PI_BIN_FILE = "{tmp}/pbins.txt"')
        else:
            lines += [f"PI_BIN_MIN = {pmin!r}", f"PI_BIN_MAX = {pmax!r}", f"PI_BIN_SIZE = {dpi!r}"]
    lines += ["OUTPUT_FORMAT = 0", "OVERWRITE = 2", "VERBOSE = F"]
    conf = tmp / "fcfc.conf"
    conf.write_text("\n".join(lines) + "\n")
    out = tmp / "out.drv"
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    pr = subprocess.run([str(exe), str(out), "-c", str(conf)], capture_output=True, text=True,
                        env=env, timeout=timeout)
    if pr.returncode != 0 or not out.exists():
        raise RuntimeError(f"reference driver failed (rc={pr.returncode}):\n{pr.stdout}\n{pr.stderr}")
    res = parse_drv(out)
    res.stdout = pr.stdout
    if not keep and workdir is None:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return res
