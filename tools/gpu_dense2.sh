#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/dense2_timeline.log; }
for rep in 1 2; do
for v in old nodense; do
  timeout 120 python tools/time_c2.py fcfc_b200/_variants/$v/libfcfc_b200.so > $O/dense2_$v.$rep.log 2>&1; el "$v: $(grep 'bt=' $O/dense2_$v.$rep.log | tr '\n' ' ')"
done
FCFC_GPU_NO_DENSE=1 timeout 120 python tools/time_c2.py fcfc_b200/libfcfc_b200.so > $O/dense2_off.$rep.log 2>&1; el "new off: $(grep 'bt=' $O/dense2_off.$rep.log | tr '\n' ' ')"
timeout 120 python tools/time_c2.py fcfc_b200/libfcfc_b200.so > $O/dense2_on.$rep.log 2>&1; el "new on: $(grep 'bt=' $O/dense2_on.$rep.log | tr '\n' ' ')"
done
el done
