#!/bin/bash
# gpurun script, round 2 / session 5: the float-speed double kernel (count_kernel_df.cuh): A/B against plain and pre-filter, GPU suite.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s5_timeline.log; }
timeout 500 python tools/time_pf.py > $O/s5_time_pf.log 2>&1; el "time_pf rc=$?"; grep -E "same|Error|error" $O/s5_time_pf.log | tee -a $O/s5_timeline.log
timeout 900 python -m pytest tests -q -m gpu > $O/s5_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s5_pytest_gpu.log)"
timeout 300 python bench.py --prec double --no-cpu > $O/s5_bench_n1_double.json 2> $O/s5_bench_n1_double.err; el "bench n1 double rc=$?"
el done
