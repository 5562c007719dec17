"""GPU (-m gpu): coordinate conversion on the device (fcfc_b200/csrc/cnvt.cu), replacing
cnvt_coord_integr (fcfc/2pt/cnvt_coord.c:321-337).

The comoving distance repeats the host's operations one by one (IEEE sqrt / division, no contraction), so it is
bit-identical; sin / cos come from the CUDA math library (<= 2 ulp) instead of libm (<= 1 ulp), so a Cartesian coordinate
may differ from the host's in its last bits (two such factors per coordinate: up to 5 ulp observed).  Bars: every coordinate within 8 ulp of the host's, the large majority
identical; and the pair counts of the real command line with FCFC_GPU_CNVT=1 equal to the stock reference's."""
import ctypes as C
import filecmp
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle

pytestmark = pytest.mark.gpu


def host_cnvt(ra, dec, z, om, ol, ok, order, gx, gw):
    lib = C.CDLL(str(oracle.build()))
    n = len(ra)
    out, dist = np.zeros(3 * n), np.zeros(n)
    dp = C.POINTER(C.c_double)
    lib.oracle_cnvt(ra.ctypes.data_as(dp), dec.ctypes.data_as(dp), z.ctypes.data_as(dp), C.c_size_t(n), C.c_double(om), C.c_double(ol),
                    C.c_double(ok), C.c_int(order), gx.ctypes.data_as(dp), gw.ctypes.data_as(dp), out.ctypes.data_as(dp), dist.ctypes.data_as(dp))
    return out.reshape(n, 3), dist


@pytest.mark.parametrize("order", [4, 7, 10, 32])
@pytest.mark.parametrize("prec", ["double", "float"])
def test_device_conversion_matches_host(gpu, order, prec):
    rng = np.random.default_rng(order)
    n = 200_000
    ra, dec, z = rng.uniform(0, 360, n), np.rad2deg(np.arcsin(rng.uniform(-1, 1, n))), rng.uniform(0.0, 2.5, n)
    z[:10] = 0.0
    x, w = np.polynomial.legendre.leggauss(order)
    pos = x > 1e-14
    gx, gw = np.ascontiguousarray(x[pos]), np.ascontiguousarray(w[pos])
    if order & 1:
        gx, gw = np.append(gx, 0.0), np.append(gw, w[np.abs(x) < 1e-14][0])
    dt = np.float32 if prec == "float" else np.float64
    a, b, c = (np.ascontiguousarray(v, dtype=dt) for v in (ra, dec, z))
    want, dist = host_cnvt(a.astype(np.float64), b.astype(np.float64), c.astype(np.float64), 0.31, 0.69, 0.0, order, gx, gw)
    L = gpu.lib()
    L.fcfc_gpu_cnvt_coord.argtypes = [C.c_void_p] * 3 + [C.c_size_t, C.c_int] + [C.c_double] * 3 + [C.c_int, C.c_void_p, C.c_void_p]
    rc = L.fcfc_gpu_cnvt_coord(a.ctypes.data, b.ctypes.data, c.ctypes.data, n, int(prec == "float"), 0.31, 0.69, 0.0, order, gx.ctypes.data, gw.ctypes.data)
    assert rc == 0
    got = np.stack([a, b, c], 1)
    want = want.astype(dt)
    ulp = np.abs(got.astype(np.float64) - want.astype(np.float64)) / np.maximum(np.spacing(np.abs(want)).astype(np.float64), 1e-300)
    same = float((got == want).mean())
    print(f"order {order} {prec}: max {ulp.max():.1f} ulp, identical coordinates {same:.4f}")
    # (coordinates that nearly vanish -- cos close to zero -- are compared on the scale of the distance)
    scale = np.maximum(np.abs(want), 1e-3 * dist[:, None]).astype(dt)
    assert (np.abs(got.astype(np.float64) - want.astype(np.float64)) <= 8 * np.spacing(scale).astype(np.float64)).all()
    assert same > (0.8 if prec == "double" else 0.99)
    # the radial distance: |x| reproduces the host's distance to rounding (the quadrature itself is bit-identical)
    rad = np.sqrt((got.astype(np.float64) ** 2).sum(1))
    np.testing.assert_allclose(rad[10:], dist[10:], rtol=(4e-7 if prec == "float" else 1e-15))


def test_fcfc_2pt_cli_with_device_conversion(tmp_path):
    gpu_bin, ref_bin = os.path.join(ROOT, "integration", "_build", "dbl", "FCFC_2PT"), os.path.join(ROOT, "oracle", "_ref", "dbl_scalar", "FCFC_2PT")
    if not (os.path.exists(gpu_bin) and os.path.exists(ref_bin)):
        pytest.skip("integration/_build or oracle/_ref not shipped")
    rng = np.random.default_rng(74)

    def radecz(n):
        return np.c_[rng.uniform(120, 180, n), np.rad2deg(np.arcsin(rng.uniform(0, 0.5, n))), rng.uniform(0.4, 0.7, n)]
    np.savetxt(tmp_path / "data.txt", radecz(20000), fmt="%.8f")
    np.savetxt(tmp_path / "rand.txt", radecz(50000), fmt="%.8f")
    for tag, exe, env in (("ref", ref_bin, {}), ("gpu", gpu_bin, {"FCFC_GPU_CNVT": "1", "FCFC_GPU_VERBOSE": "1"})):
        d = tmp_path / tag
        d.mkdir()
        (d / "fcfc.conf").write_text(f"""
CATALOG = ["{tmp_path}/data.txt","{tmp_path}/rand.txt"]
CATALOG_LABEL = [D,R]
ASCII_FORMATTER = ["%lf %lf %lf","%lf %lf %lf"]
POSITION = ["$1","$2","$3","$1","$2","$3"]
COORD_CONVERT = [T,T]
OMEGA_M = 0.31
BINNING_SCHEME = 1
PAIR_COUNT = [DD,DR,RR]
PAIR_COUNT_FILE = ["{d}/DD.bin","{d}/DR.bin","{d}/RR.bin"]
CF_ESTIMATOR = "(DD - 2*DR + RR) / RR"
CF_OUTPUT_FILE = "{d}/xi.txt"
MULTIPOLE = [0,2]
MULTIPOLE_FILE = "{d}/xil.txt"
SEP_BIN_MIN = 0
SEP_BIN_MAX = 120
SEP_BIN_SIZE = 4
MU_BIN_NUM = 40
OUTPUT_FORMAT = 0
OVERWRITE = 2
VERBOSE = F
""")
        r = subprocess.run([exe, "-c", str(d / "fcfc.conf")], cwd=str(d), capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="8", **env), timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        if tag == "gpu":
            assert "converted on the device" in r.stderr
    # ~10^9 pairs, positions equal to a few 1e-16 relative: the counts are identical (a pair would have to sit within
    # 1e-13 of a bin edge to move)
    for f in ("DD.bin", "DR.bin", "RR.bin", "xi.txt", "xil.txt"):
        assert filecmp.cmp(tmp_path / "ref" / f, tmp_path / "gpu" / f, shallow=False), f
