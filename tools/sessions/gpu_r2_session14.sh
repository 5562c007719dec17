#!/bin/bash
# gpurun script: rotating stack columns in the one-point-per-step pair loop (weighted, (s_perp,pi), survey, plain double): A/B against -DFCFC_ROTATE=0.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s14_timeline.log; }
for v in rot norot; do
  if [ $v = norot ]; then export FCFC_B200_LIB=$PWD/fcfc_b200/_variants/norot/libfcfc_b200.so; else unset FCFC_B200_LIB; fi
  timeout 300 python tools/time_wt.py > $O/s14_wt_$v.log 2>&1; el "wt $v rc=$?"
  timeout 300 python tools/time_survey.py 200000 2000000 float > $O/s14_svy_float_$v.log 2>&1; el "survey float $v rc=$?"
  timeout 300 python tools/time_survey.py 200000 2000000 double > $O/s14_svy_double_$v.log 2>&1; el "survey double $v rc=$?"
  FCFC_GPU_TUNE="no_df=1,no_prefilter=1" timeout 300 python tools/time_wt.py > $O/s14_wt_plain_$v.log 2>&1; el "wt plain double $v rc=$?"
done
unset FCFC_B200_LIB
for f in wt svy_float svy_double wt_plain; do echo "== $f"; paste -d'\n' $O/s14_${f}_rot.log $O/s14_${f}_norot.log; done | tee -a $O/s14_timeline.log
timeout 900 python -m pytest tests -q -m gpu -x > $O/s14_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s14_pytest_gpu.log)"
el done
