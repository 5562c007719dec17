#!/bin/bash
# gpurun script: clustered 10^7-point mock against the reference's counts (tests/golden/fullsize_c4.npz): GPU tests + bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s31_timeline.log; }
timeout 600 python -m pytest tests/test_fullsize_golden.py -q -m gpu -k "c4]" -s > $O/s31_pytest.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s31_pytest.log)"; grep "c4 arith" $O/s31_pytest.log | tee -a $O/s31_timeline.log
el done
