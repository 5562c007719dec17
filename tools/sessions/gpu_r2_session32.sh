#!/bin/bash
# gpurun script: survey tests after the slack change of the cylinder classification (pre-filter kernel).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s32_timeline.log; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_golden.py -q -m gpu -k "survey or svy" > $O/s32_pytest.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s32_pytest.log)"
el done
