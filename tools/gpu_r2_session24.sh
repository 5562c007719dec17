#!/bin/bash
# gpurun script: survey (s_perp,pi) variants of the pre-filter kernel at 28 / 32 warps.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for v in main pf32; do
  if [ $v = main ]; then unset FCFC_B200_LIB; else export FCFC_B200_LIB=$PWD/fcfc_b200/_variants/$v/libfcfc_b200.so; fi
  FCFC_TS_BINTYPES=2 FCFC_TS_WEIGHTED_ONLY=1 timeout 300 python tools/time_survey.py 200000 2000000 double 2>&1 | cut -c1-220
done
