"""GPU (-m gpu): the live peak measurements behind bench.py's roofline (fcfc_b200/csrc/peaks.cu) return plausible numbers
for a B200: FP32 issue 128 lanes/clk/SM (148 SMs at <= 2.1 GHz), FP64 issue, shared-memory histogram increments (a few
lanes per clock and SM: DESIGN.md section 4 quotes 8-9 from tools/microbench.cu)."""
import pytest

pytestmark = pytest.mark.gpu


def test_peak_measurements_are_plausible(gpu):
    fp32, mhz = gpu.measure_fp32_peak()
    assert 1.0e13 < fp32 < 4.5e13, fp32
    assert 800 < mhz < 2300, mhz
    fp64 = gpu.measure_fp64_peak()
    assert 0.2 * fp32 < fp64 < 0.75 * fp32, (fp32, fp64)
    atoms = gpu.measure_smem_atomic_peak()
    per_clk_sm = 128.0 * atoms / fp32
    assert 2.0 < per_clk_sm < 40.0, per_clk_sm
