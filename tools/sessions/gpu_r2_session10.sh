#!/bin/bash
# gpurun script: rotating stack columns: parity, then A/B timings (qkeep scan) on the bench workload.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s10_timeline.log; }
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/s10_pytest_parity.log 2>&1; el "pytest parity rc=$?: $(tail -1 $O/s10_pytest_parity.log)"
timeout 600 python tools/time_opts.py "no_classify=1" "no_classify=0" "qkeep=8" "qkeep=10" "qkeep=12" "qkeep=4" > $O/s10_time_opts.log 2>&1; el "time_opts rc=$?"
cat $O/s10_time_opts.log | tee -a $O/s10_timeline.log
el done
