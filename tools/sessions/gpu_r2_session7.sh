#!/bin/bash
# gpurun script: default bench line (with its double-precision part), configs[2] survey bench, ncu captures of the float-speed
# double kernel and of the float kernel at 2x10^6 points.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s7_timeline.log; }
timeout 300 python bench.py > $O/s7_bench_n1.json 2> $O/s7_bench_n1.err; el "bench rc=$?"
timeout 600 python bench.py --workload c3_svy_spi_wt_2e6_2e7 --steps 2 --warmup 1 > $O/s7_bench_c3.json 2> $O/s7_bench_c3.err; el "bench c3 rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:count_kernel_df -s 1 -c 1 -f -o $O/ncu_df_2e6 python tools/prof_pf.py - > $O/s7_ncu_df.log 2>&1; el "ncu df rc=$?"
python tools/ncu_summary.py $O/ncu_df_2e6.ncu-rep 0.01 > $O/ncu_df_2e6_summary.txt 2>&1; el "summary df rc=$?"
timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:count_kernel -s 1 -c 1 --csv --log-file $O/s7_ncu_traffic_c2.csv python tools/prof_one.py 1e7 2000 1 float 1 2 > /dev/null 2>&1; el "traffic rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s7_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-double > $O/s7_ncu_bench_under_ncu.log 2>&1; el "launch list rc=$?"
el done
