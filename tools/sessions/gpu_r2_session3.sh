#!/bin/bash
# gpurun script, round 2 / session 3: the pre-filter with bit masks (pf2): A/B against the plain double kernels, GPU suite.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s3_timeline.log; }
timeout 400 python tools/time_pf.py > $O/s3_time_pf.log 2>&1; el "time_pf rc=$?"; grep same $O/s3_time_pf.log | tee -a $O/s3_timeline.log
timeout 900 python -m pytest tests -q -m gpu -x > $O/s3_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s3_pytest_gpu.log)"
el done
