#!/bin/bash
# gpurun script: where the time of the pre-filter kernel goes (diagnostics build + one ncu capture at 2x10^6 points).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s4_timeline.log; }
timeout 120 python tools/prof_pf.py fcfc_b200/_variants/pfstats/libfcfc_b200.so > $O/s4_pfstats.log 2>&1; el "pfstats rc=$?"; cat $O/s4_pfstats.log | tee -a $O/s4_timeline.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:count_kernel_pf -s 1 -c 1 -f -o $O/ncu_pf_2e6 python tools/prof_pf.py - > $O/s4_ncu.log 2>&1; el "ncu rc=$?"
python tools/ncu_summary.py $O/ncu_pf_2e6.ncu-rep 0.01 > $O/ncu_pf_2e6_summary.txt 2>&1; el "summary rc=$?"
el done
