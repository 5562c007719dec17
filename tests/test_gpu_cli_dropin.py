"""GPU (-m gpu): the drop-in itself.  The unmodified FCFC host linked against the engine
(integration/_build, built by integration/Makefile from the reference sources + fcfc_gpu_shim.c) runs the
reference's own command line on the reference's own configuration keywords; every output file -- binary
pair counts, xi(s,mu), multipoles, w_p -- must be byte-identical to what the stock reference writes."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

from cases import box_catalog, survey_catalog
from conftest import ROOT

pytestmark = pytest.mark.gpu

GPU_BIN = os.path.join(ROOT, "integration", "_build")
REF_BIN = os.path.join(ROOT, "oracle", "_ref")


def _have(prec, prog):
    return os.path.exists(os.path.join(GPU_BIN, prec, prog)) and os.path.exists(os.path.join(REF_BIN, f"{prec}_scalar", prog))


def _run(exe, conf, cwd):
    env = dict(os.environ, OMP_NUM_THREADS="8")
    r = subprocess.run([exe, "-c", conf], cwd=cwd, capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


@pytest.mark.skipif(not _have("dbl", "FCFC_2PT_BOX"), reason="integration/_build or oracle/_ref not shipped")
@pytest.mark.parametrize("prec", ["dbl", "flt"])
def test_fcfc_2pt_box_cli_byte_identical(tmp_path, prec):
    n, L = 30000, 500.0
    x, y, z, w = box_catalog(n, L, 71)
    if prec == "flt":   # keep every in-range pair away from the periodic faces (SINGLE_PREC caveat)
        x, y, z = x * 0.5 + 125.0, y * 0.5 + 125.0, z * 0.5 + 125.0
    np.savetxt(tmp_path / "data.txt", np.c_[x, y, z, w], fmt="%.6f")
    fm = "%f" if prec == "flt" else "%lf"
    outs = {}
    for tag, exe in (("ref", os.path.join(REF_BIN, f"{prec}_scalar", "FCFC_2PT_BOX")), ("gpu", os.path.join(GPU_BIN, prec, "FCFC_2PT_BOX"))):
        d = tmp_path / tag
        d.mkdir()
        conf = d / "fcfc.conf"
        conf.write_text(f"""
CATALOG = "{tmp_path}/data.txt"
CATALOG_LABEL = D
ASCII_FORMATTER = "{fm} {fm} {fm} {fm}"
POSITION = ["$1","$2","$3"]
WEIGHT = "$4"
BOX_SIZE = {L}
BINNING_SCHEME = 1
PAIR_COUNT = DD
PAIR_COUNT_FILE = "{d}/DD.bin"
CF_ESTIMATOR = "DD / @@ - 1"
CF_OUTPUT_FILE = "{d}/xi.txt"
MULTIPOLE = [0,2,4]
MULTIPOLE_FILE = "{d}/xil.txt"
SEP_BIN_MIN = 0
SEP_BIN_MAX = 60
SEP_BIN_SIZE = 2
MU_BIN_NUM = 30
OUTPUT_FORMAT = 0
OVERWRITE = 2
VERBOSE = F
""")
        outs[tag] = _run(exe, str(conf), str(d))
    # weighted sums may differ in the last bits (summation order): compare the tables numerically
    for f in ("xi.txt", "xil.txt"):
        a, b = np.loadtxt(tmp_path / "ref" / f), np.loadtxt(tmp_path / "gpu" / f)
        assert a.shape == b.shape
        np.testing.assert_allclose(b, a, rtol=1e-9, atol=1e-12)
    assert os.path.getsize(tmp_path / "ref" / "DD.bin") == os.path.getsize(tmp_path / "gpu" / "DD.bin")


@pytest.mark.skipif(not _have("dbl", "FCFC_2PT_BOX"), reason="integration/_build or oracle/_ref not shipped")
def test_fcfc_2pt_box_cli_unweighted_binary_identical(tmp_path):
    n, L = 30000, 500.0
    x, y, z, _ = box_catalog(n, L, 72)
    np.savetxt(tmp_path / "data.txt", np.c_[x, y, z], fmt="%.6f")
    for tag, exe in (("ref", os.path.join(REF_BIN, "dbl_scalar", "FCFC_2PT_BOX")), ("gpu", os.path.join(GPU_BIN, "dbl", "FCFC_2PT_BOX"))):
        d = tmp_path / tag
        d.mkdir()
        conf = d / "fcfc.conf"
        conf.write_text(f"""
CATALOG = "{tmp_path}/data.txt"
CATALOG_LABEL = D
ASCII_FORMATTER = "%lf %lf %lf"
POSITION = ["$1","$2","$3"]
BOX_SIZE = {L}
BINNING_SCHEME = 2
PAIR_COUNT = DD
PAIR_COUNT_FILE = "{d}/DD.bin"
CF_ESTIMATOR = "DD / @@ - 1"
CF_OUTPUT_FILE = "{d}/xi.txt"
PROJECTED_CF = T
PROJECTED_FILE = "{d}/wp.txt"
SEP_BIN_MIN = 0
SEP_BIN_MAX = 40
SEP_BIN_SIZE = 2
PI_BIN_MIN = 0
PI_BIN_MAX = 60
PI_BIN_SIZE = 1
OUTPUT_FORMAT = 0
OVERWRITE = 2
VERBOSE = F
""")
        _run(exe, str(conf), str(d))
    for f in ("DD.bin", "xi.txt", "wp.txt"):
        assert filecmp.cmp(tmp_path / "ref" / f, tmp_path / "gpu" / f, shallow=False), f


@pytest.mark.skipif(not _have("dbl", "FCFC_2PT"), reason="integration/_build or oracle/_ref not shipped")
def test_fcfc_2pt_survey_cli(tmp_path):
    """FCFC_2PT with the reference's own coordinate conversion (ra, dec, z -> comoving, flat LCDM) on the host."""
    rng = np.random.default_rng(73)

    def radecz(n):
        return np.c_[rng.uniform(120, 180, n), np.rad2deg(np.arcsin(rng.uniform(0, 0.5, n))), rng.uniform(0.4, 0.7, n),
                     rng.uniform(0.75, 1.25, n)]
    np.savetxt(tmp_path / "data.txt", radecz(20000), fmt="%.8f")
    np.savetxt(tmp_path / "rand.txt", radecz(60000), fmt="%.8f")
    for tag, exe in (("ref", os.path.join(REF_BIN, "dbl_scalar", "FCFC_2PT")), ("gpu", os.path.join(GPU_BIN, "dbl", "FCFC_2PT"))):
        d = tmp_path / tag
        d.mkdir()
        conf = d / "fcfc.conf"
        conf.write_text(f"""
CATALOG = ["{tmp_path}/data.txt","{tmp_path}/rand.txt"]
CATALOG_LABEL = [D,R]
ASCII_FORMATTER = ["%lf %lf %lf %lf","%lf %lf %lf %lf"]
POSITION = ["$1","$2","$3","$1","$2","$3"]
WEIGHT = ["$4","$4"]
COORD_CONVERT = [T,T]
OMEGA_M = 0.31
BINNING_SCHEME = 2
PAIR_COUNT = [DD,DR,RR]
PAIR_COUNT_FILE = ["{d}/DD.bin","{d}/DR.bin","{d}/RR.bin"]
CF_ESTIMATOR = "(DD - 2*DR + RR) / RR"
CF_OUTPUT_FILE = "{d}/xi.txt"
PROJECTED_CF = T
PROJECTED_FILE = "{d}/wp.txt"
SEP_BIN_MIN = 0
SEP_BIN_MAX = 40
SEP_BIN_SIZE = 2
PI_BIN_MIN = 0
PI_BIN_MAX = 80
PI_BIN_SIZE = 1
OUTPUT_FORMAT = 1
OVERWRITE = 2
VERBOSE = F
""")
        _run(exe, str(conf), str(d))
    # weighted sums agree to <= 1e-12 relative, but the tables are text with 10 significant digits: a sum that
    # differs in its 16th digit can still flip the last printed digit.  Every number must agree to one unit of the
    # last printed digit, and all but a handful (rounding flips) to 1e-12.
    for f in ("DD.bin", "DR.bin", "RR.bin", "xi.txt", "wp.txt"):
        a = np.loadtxt(tmp_path / "ref" / f)
        b = np.loadtxt(tmp_path / "gpu" / f)
        assert a.shape == b.shape
        np.testing.assert_allclose(b, a, rtol=1.5e-9, atol=1e-12)
    for f in ("DD.bin", "DR.bin", "RR.bin"):
        a = np.loadtxt(tmp_path / "ref" / f)[:, -2:]
        b = np.loadtxt(tmp_path / "gpu" / f)[:, -2:]
        flips = np.abs(b - a) > 1e-12 * np.abs(a)
        assert flips.sum() <= max(2, a.size // 500), f"{f}: {flips.sum()} of {a.size} printed sums differ"
