#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/extra_timeline.log; }
timeout 300 python bench.py --workload c4_box_smu_clustered_1e7 --steps 3 > $O/extra_bench_clustered.json 2> $O/extra_bench_clustered.err; el "clustered rc=$?"
timeout 400 python bench.py --workload c3_svy_spi_wt_2e6_2e7 --steps 2 > $O/extra_bench_survey.json 2> $O/extra_bench_survey.err; el "survey rc=$?"
tail -5 $O/extra_bench_survey.err
el done
