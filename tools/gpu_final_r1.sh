#!/bin/bash
# One gpurun call: time the experimental kernel builds (fcfc_b200/_variants/*), pick the fastest one whose counts
# are identical to the default build on every check workload, run the GPU test suite and bench.py with it, then the
# other timings of the round.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/final_timeline.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/final_smi.log 2>&1
BASE=fcfc_b200/libfcfc_b200.so
cp $BASE /tmp/base_lib.so
el "variant checks"
timeout 240 python tools/variant_check.py $BASE > $O/variant_base.log 2>&1; el "base rc=$?"
LOGS="$O/variant_base.log"
for v in ${VARIANTS:-flag2 flag1}; do
  if [ -f fcfc_b200/_variants/$v/libfcfc_b200.so ]; then
    timeout 240 python tools/variant_check.py fcfc_b200/_variants/$v/libfcfc_b200.so > $O/variant_$v.log 2>&1; el "$v rc=$?"
    LOGS="$LOGS $O/variant_$v.log"
  fi
done
BEST=$(python tools/pick_variant.py $LOGS 2> $O/variant_pick.log); cat $O/variant_pick.log
el "picked $BEST"
W=$(basename $BEST .log); W=${W#variant_}
echo "$W" > $O/variant_winner.txt
if [ "$W" != "base" ]; then cp fcfc_b200/_variants/$W/libfcfc_b200.so $BASE; fi
el "pytest -m gpu with $W"
timeout 420 python -m pytest tests -x -q -m gpu > $O/final_pytest_gpu.log 2>&1; RC=$?; el "pytest rc=$RC"; tail -3 $O/final_pytest_gpu.log
if [ $RC -ne 0 ] && [ "$W" != "base" ]; then
  el "winner failed the suite: back to the default build"
  cp /tmp/base_lib.so $BASE; echo "base (after $W failed pytest)" > $O/variant_winner.txt
  timeout 420 python -m pytest tests -x -q -m gpu > $O/final_pytest_gpu_base.log 2>&1; el "pytest base rc=$?"; tail -3 $O/final_pytest_gpu_base.log
fi
el "bench"
timeout 300 python bench.py > $O/final_bench.json 2> $O/final_bench.err; el "bench rc=$?"; cat $O/final_bench.json
el "extras"
timeout 120 python tools/time_clustered.py 2e6 1169.6 > $O/final_clustered.log 2>&1; el "clustered rc=$?"
FCFC_TS_BINTYPES=2 FCFC_TS_WEIGHTED_ONLY=1 timeout 150 python tools/time_survey.py 2e6 2e7 double > $O/final_survey_c3.log 2>&1; el "survey C3 rc=$?"
timeout 100 python tools/time_c2.py $BASE > $O/final_time_c2.log 2>&1; el "time_c2 rc=$?"
el "done"
