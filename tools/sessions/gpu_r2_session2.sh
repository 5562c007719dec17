#!/bin/bash
# gpurun script, round 2 / session 2: in-place binning (option inplace) A/B, correctness of the GPU suite with it on.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s2_timeline.log; }
timeout 300 python tools/time_opts.py "inplace=0" "inplace=1" "inplace=1,k=6" "inplace=1,k=7" "inplace=0,k=6" > $O/s2_time_opts.log 2>&1; el "time_opts rc=$?"; cat $O/s2_time_opts.log | tee -a $O/s2_timeline.log
FCFC_GPU_TUNE="inplace=1" timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_golden.py -q -m gpu -x > $O/s2_pytest_inplace.log 2>&1; el "pytest inplace rc=$?: $(tail -1 $O/s2_pytest_inplace.log)"
el done
