#!/bin/bash
# gpurun script: ncu full capture (source counters) of the classified-staging kernel at 2x10^6 points.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s9_timeline.log; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:count_kernel_cl -s 1 -c 1 -f -o $O/ncu_cl_2e6 python tools/prof_one.py 2e6 1169.6 1 float 1 2 > $O/s9_ncu_full.log 2>&1; el "full capture rc=$?"
python tools/ncu_summary.py $O/ncu_cl_2e6.ncu-rep 0.005 > $O/ncu_cl_2e6_summary.txt 2>&1; el "summary rc=$?"
el done
