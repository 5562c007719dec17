#!/bin/bash
# gpurun script: classified staging (count_kernel_cl.cuh): parity first, then A/B timings on the bench workload.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s8_timeline.log; }
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/s8_pytest_parity.log 2>&1; el "pytest parity rc=$?: $(tail -1 $O/s8_pytest_parity.log)"
timeout 600 python tools/time_opts.py "no_classify=1" "no_classify=0" "no_classify=0,k=4" "no_classify=0,k=6" "no_classify=0,k=7" > $O/s8_time_opts.log 2>&1; el "time_opts rc=$?"
cat $O/s8_time_opts.log | tee -a $O/s8_timeline.log
timeout 900 python -m pytest tests/test_fullsize_golden.py -q -m gpu -x > $O/s8_pytest_fullsize.log 2>&1; el "pytest fullsize rc=$?: $(tail -1 $O/s8_pytest_fullsize.log)"
el done
