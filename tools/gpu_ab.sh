#!/bin/bash
# gpurun script: A/B of the default build with the dense-cell path off / on (timings, count digests), GPU suite.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/ab_timeline.log; }
LIB=fcfc_b200/libfcfc_b200.so
for rep in 1 2; do
FCFC_GPU_NO_DENSE=1 timeout 120 python tools/time_c2.py $LIB > $O/ab_off.$rep.log 2>&1; el "off: $(grep 'bt=' $O/ab_off.$rep.log | tr '\n' ' ')"
timeout 120 python tools/time_c2.py $LIB > $O/ab_on.$rep.log 2>&1; el "on: $(grep 'bt=' $O/ab_on.$rep.log | tr '\n' ' ')"
done
FCFC_GPU_NO_DENSE=1 timeout 240 python tools/variant_check.py $LIB > $O/ab_vc_off.log 2>&1; el "vc off rc=$?"
timeout 240 python tools/variant_check.py $LIB > $O/ab_vc_on.log 2>&1; el "vc on rc=$?"
python tools/pick_variant.py $O/ab_vc_off.log $O/ab_vc_on.log 2> $O/ab_pick.log; cat $O/ab_pick.log
if [ "${1:-}" = "tests" ]; then timeout 420 python -m pytest tests -x -q -m gpu > $O/ab_pytest_gpu.log 2>&1; el "pytest rc=$?"; tail -2 $O/ab_pytest_gpu.log; fi
el done
