#!/bin/bash
# gpurun script: ncu capture of the survey (s_perp,pi) weighted double kernel (float pre-filter + exact pass), RR of 2x10^6 randoms.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s15_timeline.log; }
FCFC_TS_BINTYPES=2 FCFC_TS_WEIGHTED_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_kernel_pf -s 5 -c 1 -f -o $O/ncu_pf_svy_spi python tools/time_survey.py 200000 2000000 double > $O/s15_ncu.log 2>&1; el "capture rc=$?"
python tools/ncu_summary.py $O/ncu_pf_svy_spi.ncu-rep 0.01 > $O/ncu_pf_svy_spi_summary.txt 2>&1; el "summary rc=$?"
el done
