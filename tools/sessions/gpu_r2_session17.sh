#!/bin/bash
# gpurun script: ncu evidence for the kernels of this round: classified-staging float kernel (full capture at 2x10^6 points,
# DRAM traffic of one C2 launch, launch list of a bench run) and the survey pre-filter kernel after classification.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s17_timeline.log; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:count_kernel_cl -s 1 -c 1 -f -o $O/ncu_cl_final_2e6 python tools/prof_one.py 2e6 1169.6 1 float 1 2 > $O/s17_ncu_cl.log 2>&1; el "cl capture rc=$?"
python tools/ncu_summary.py $O/ncu_cl_final_2e6.ncu-rep 0.01 > $O/ncu_cl_final_2e6_summary.txt 2>&1; el "summary rc=$?"
FCFC_TS_BINTYPES=2 FCFC_TS_WEIGHTED_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_kernel_pf -s 5 -c 1 -f -o $O/ncu_pf_svy_spi_after python tools/time_survey.py 200000 2000000 double > $O/s17_ncu_pf.log 2>&1; el "pf capture rc=$?"
python tools/ncu_summary.py $O/ncu_pf_svy_spi_after.ncu-rep 0.01 > $O/ncu_pf_svy_spi_after_summary.txt 2>&1; el "summary rc=$?"
timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:count_kernel -s 1 -c 1 --csv --log-file $O/s17_ncu_traffic_c2.csv python tools/prof_one.py 1e7 2000 1 float 1 2 > /dev/null 2>&1; el "traffic rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s17_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-double > $O/s17_bench_under_ncu.log 2>&1; el "launch list rc=$?"
el done
