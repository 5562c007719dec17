#!/bin/bash
# gpurun script: pre-filter kernel at 24 warps per block: GPU suite, configs[2] bench, survey timings.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s25_timeline.log; }
timeout 900 python -m pytest tests -q -m gpu -x > $O/s25_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s25_pytest_gpu.log)"
timeout 300 python tools/time_survey.py 200000 2000000 double > $O/s25_svy.log 2>&1; el "survey rc=$?"; cat $O/s25_svy.log | cut -c1-220 | tee -a $O/s25_timeline.log
timeout 600 python bench.py --workload c3_svy_spi_wt_2e6_2e7 --steps 2 --warmup 1 > $O/s25_bench_c3.json 2> $O/s25_bench_c3.err; el "bench c3 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/s25_bench_c3.json').read().strip().splitlines()[-1]); print('c3 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('max_rel_diff_vs_resident'))" | tee -a $O/s25_timeline.log
el done
