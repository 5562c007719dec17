#!/bin/bash
# gpurun script: warps per block of the pre-filter kernel (16 / 20 / 24) on the survey (s_perp,pi) weighted counts.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s24_timeline.log; }
for v in main pf16 pf24; do
  if [ $v = main ]; then unset FCFC_B200_LIB; else export FCFC_B200_LIB=$PWD/fcfc_b200/_variants/$v/libfcfc_b200.so; fi
  FCFC_TS_BINTYPES=2,1 FCFC_TS_WEIGHTED_ONLY=1 timeout 300 python tools/time_survey.py 200000 2000000 double > $O/s24_svy_$v.log 2>&1; el "survey $v rc=$?"; cat $O/s24_svy_$v.log | cut -c1-220 | tee -a $O/s24_timeline.log
done
el done
