"""CPU: the C-ABI shared library loads, exports every symbol include/fcfc_gpu.h declares, and refuses to
compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import fcfc_b200 as F
from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fcfc_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fcfc_gpu_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_seam():
    syms = declared_symbols()
    for s in ("fcfc_gpu_init", "fcfc_gpu_catalog_create", "fcfc_gpu_catalog_destroy", "fcfc_gpu_count",
              "fcfc_gpu_count_partial", "fcfc_gpu_last_error", "fcfc_gpu_bins_create"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    L = F.lib()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    assert L.fcfc_gpu_abi_version() == 1


def test_struct_layout_matches_header():
    # 12 int32 + 5 pointers + 2 uint64 + 3 double
    assert ctypes.sizeof(F.api._CBins) == 12 * 4 + 5 * 8 + 2 * 8 + 3 * 8
    # 2 uint64 + 3 double + kernel_launches + ncell[3] + nitem + dense_rows + prefilter + classified (int32 each) + uint64
    assert ctypes.sizeof(F.api._CStats) == 2 * 8 + 3 * 8 + 4 + 3 * 4 + 4 * 4 + 8


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_no_cpu_fallback_without_device():
    with pytest.raises(F.FcfcGpuError, match="no CPU fallback"):
        F.init()
    b = F.Bins(periodic=True, bintype=0, smax=10.0, ds=1.0, box=100.0)
    with pytest.raises(F.FcfcGpuError):
        F.Catalog([1.0], [2.0], [3.0], bins=b)
    with pytest.raises(F.FcfcGpuError):
        F.CatalogStream(bins=b)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under fcfc_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fcfc_b200")):
        if "_build" in dirpath or "_gen" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the reference's oracle", ""), f"{f} mentions the oracle"
