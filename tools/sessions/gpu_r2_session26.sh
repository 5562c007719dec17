#!/bin/bash
# gpurun script: survey isotropic float through count_kernel_cl, half-box reach test, full GPU suite, survey float timings.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s26_timeline.log; }
timeout 900 python -m pytest tests -q -m gpu -x > $O/s26_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -3 $O/s26_pytest_gpu.log | tr '\n' ' ')"
FCFC_TS_BINTYPES=0 timeout 300 python tools/time_survey.py 200000 2000000 float > $O/s26_svy_float.log 2>&1; el "survey float rc=$?"; cat $O/s26_svy_float.log | cut -c1-220 | tee -a $O/s26_timeline.log
el done
