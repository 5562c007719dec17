#!/bin/bash
# gpurun script: balanced exact pass of the pre-filter kernel: GPU suite, survey / weighted timings, configs[2] bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s16_timeline.log; }
timeout 900 python -m pytest tests -q -m gpu -x > $O/s16_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s16_pytest_gpu.log)"
timeout 300 python tools/time_survey.py 200000 2000000 double > $O/s16_svy_double.log 2>&1; el "survey double rc=$?"; cat $O/s16_svy_double.log | tee -a $O/s16_timeline.log
timeout 300 python tools/time_wt.py > $O/s16_wt.log 2>&1; el "wt rc=$?"; grep double $O/s16_wt.log | tee -a $O/s16_timeline.log
timeout 600 python bench.py --workload c3_svy_spi_wt_2e6_2e7 --steps 2 --warmup 1 > $O/s16_bench_c3.json 2> $O/s16_bench_c3.err; el "bench c3 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/s16_bench_c3.json').read().strip().splitlines()[-1]); print('c3 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" | tee -a $O/s16_timeline.log
el done
