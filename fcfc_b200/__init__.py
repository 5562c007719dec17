"""fcfc_b200 -- B200-native (sm_100a) pair-counting engine: a drop-in for the counting step of
cheng-zhao/FCFC (FCFC_2PT_BOX / FCFC_2PT).  See DESIGN.md, INTEGRATION.md and include/fcfc_gpu.h."""
from .api import (ARITH_FMA, ARITH_SCALAR, BIN_ISO, BIN_SMU, BIN_SPI, Bins, Catalog, CatalogStream, FcfcGpuError, count_pairs,
                  init, last_error, lib, measure_fp32_peak, measure_fp64_peak, measure_smem_atomic_peak, n_linear_bins, set_option, stats)

__all__ = ["ARITH_FMA", "ARITH_SCALAR", "BIN_ISO", "BIN_SMU", "BIN_SPI", "Bins", "Catalog", "CatalogStream", "FcfcGpuError",
           "count_pairs", "init", "last_error", "lib", "measure_fp32_peak", "measure_fp64_peak", "measure_smem_atomic_peak", "n_linear_bins", "set_option", "stats"]
