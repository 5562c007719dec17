#!/bin/bash
# gpurun script: new GPU tests of the classified paths; tiles of 96 points for the double-precision float-speed kernel (A/B).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s21_timeline.log; }
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "classif" > $O/s21_pytest_new.log 2>&1; el "pytest new rc=$?: $(tail -1 $O/s21_pytest_new.log)"
timeout 600 python tools/time_c2_double.py fcfc_b200/libfcfc_b200.so fcfc_b200/_variants/dfr3/libfcfc_b200.so > $O/s21_double.log 2>&1; el "double rc=$?"; cat $O/s21_double.log | tee -a $O/s21_timeline.log
el done
