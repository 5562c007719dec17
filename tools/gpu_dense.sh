#!/bin/bash
# gpurun script: dense-cell path on/off (FCFC_GPU_NO_DENSE), parity digests, cell-size scan, GPU test suite.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/dense_timeline.log; }
LIB=fcfc_b200/libfcfc_b200.so
FCFC_GPU_NO_DENSE=1 timeout 240 python tools/variant_check.py $LIB > $O/dense_off.log 2>&1; el "off rc=$?"
timeout 240 python tools/variant_check.py $LIB > $O/dense_on.log 2>&1; el "on rc=$?"
python tools/pick_variant.py $O/dense_off.log $O/dense_on.log 2> $O/dense_pick.log; cat $O/dense_pick.log
for k in 4 6 7; do
  FCFC_GPU_K=$k timeout 120 python tools/time_c2.py $LIB > $O/dense_k$k.log 2>&1; el "k=$k rc=$?: $(grep 'bt=1' $O/dense_k$k.log)"
done
for k in 5 6; do
  FCFC_GPU_NO_DENSE=1 FCFC_GPU_K=$k timeout 120 python tools/time_c2.py $LIB > $O/dense_off_k$k.log 2>&1; el "off k=$k rc=$?: $(grep 'bt=1' $O/dense_off_k$k.log)"
done
timeout 420 python -m pytest tests -x -q -m gpu > $O/dense_pytest_gpu.log 2>&1; el "pytest rc=$?"; tail -3 $O/dense_pytest_gpu.log
el done
