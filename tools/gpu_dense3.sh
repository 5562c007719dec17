#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/dense3_timeline.log; }
for v in old; do
  timeout 120 python tools/time_c2.py fcfc_b200/_variants/$v/libfcfc_b200.so > $O/dense3_$v.log 2>&1; el "$v: $(grep 'bt=' $O/dense3_$v.log | tr '\n' ' ')"
done
timeout 120 python tools/time_c2.py fcfc_b200/libfcfc_b200.so > $O/dense3_on.log 2>&1; el "new on: $(grep 'bt=' $O/dense3_on.log | tr '\n' ' ')"
FCFC_GPU_NO_DENSE=1 timeout 240 python tools/variant_check.py fcfc_b200/libfcfc_b200.so > $O/dense3_vc_off.log 2>&1; el "vc off rc=$?"
timeout 240 python tools/variant_check.py fcfc_b200/libfcfc_b200.so > $O/dense3_vc_on.log 2>&1; el "vc on rc=$?"
python tools/pick_variant.py $O/dense3_vc_off.log $O/dense3_vc_on.log 2> $O/dense3_pick.log; cat $O/dense3_pick.log
timeout 420 python -m pytest tests -x -q -m gpu > $O/dense3_pytest_gpu.log 2>&1; el "pytest rc=$?"; tail -3 $O/dense3_pytest_gpu.log
timeout 300 python bench.py > $O/dense3_bench.json 2> $O/dense3_bench.err; el "bench rc=$?"
timeout 100 python tools/time_clustered.py 2e6 1169.6 > $O/dense3_clustered.log 2>&1; el "clustered rc=$?"; cat $O/dense3_clustered.log
el done
