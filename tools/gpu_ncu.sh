#!/bin/bash
# gpurun script: ncu evidence for the current kernel (launch list of a bench run, DRAM traffic of one C2 launch,
# one full capture at 2x10^6 points with source counters).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/ncu_timeline.log; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_bench_under_ncu.log 2>&1; el "launch list rc=$?"
timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:count_kernel -s 1 -c 1 --csv --log-file $O/ncu_traffic_c2.csv python tools/prof_one.py 1e7 2000 1 float 1 2 > /dev/null 2>&1; el "traffic rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 1 -c 1 -f -o $O/ncu_count_kernel_2e6 python tools/prof_one.py 2e6 1169.6 1 float 1 2 > $O/ncu_full.log 2>&1; el "full capture rc=$?"
python tools/ncu_summary.py $O/ncu_count_kernel_2e6.ncu-rep 0.01 > $O/ncu_count_kernel_2e6_summary.txt 2>&1; el "summary rc=$?"
ls -la $O/*.ncu-rep
el done
