#!/bin/bash
# gpurun script: dense blocks without range test / stack (dense_block): parity, timing, ncu capture.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s11_timeline.log; }
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/s11_pytest_parity.log 2>&1; el "pytest parity rc=$?: $(tail -1 $O/s11_pytest_parity.log)"
timeout 600 python tools/time_opts.py "no_classify=1" "no_classify=0" "qkeep=8" > $O/s11_time_opts.log 2>&1; el "time_opts rc=$?"
cat $O/s11_time_opts.log | tee -a $O/s11_timeline.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:count_kernel_cl -s 1 -c 1 -f -o $O/ncu_cl2_2e6 python tools/prof_one.py 2e6 1169.6 1 float 1 2 > $O/s11_ncu_full.log 2>&1; el "full capture rc=$?"
python tools/ncu_summary.py $O/ncu_cl2_2e6.ncu-rep 0.005 > $O/ncu_cl2_2e6_summary.txt 2>&1; el "summary rc=$?"
el done
