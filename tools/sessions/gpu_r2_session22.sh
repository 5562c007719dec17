#!/bin/bash
# gpurun script: copies of the weighted shared-memory histogram in the pre-filter and float-speed kernels: tests + timings.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s22_timeline.log; }
timeout 900 python -m pytest tests -q -m gpu -x > $O/s22_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s22_pytest_gpu.log)"
for v in copies nocopies; do
  if [ $v = nocopies ]; then export FCFC_GPU_TUNE="no_hist_copies=1"; else unset FCFC_GPU_TUNE; fi
  FCFC_TS_BINTYPES=2,1 FCFC_TS_WEIGHTED_ONLY=1 timeout 300 python tools/time_survey.py 200000 2000000 double > $O/s22_svy_$v.log 2>&1; el "survey $v rc=$?"; cat $O/s22_svy_$v.log | cut -c1-200 | tee -a $O/s22_timeline.log
  timeout 300 python tools/time_wt.py > $O/s22_wt_$v.log 2>&1; el "wt $v rc=$?"; grep "double.*True" $O/s22_wt_$v.log | tee -a $O/s22_timeline.log
done
unset FCFC_GPU_TUNE
timeout 600 python bench.py --workload c3_svy_spi_wt_2e6_2e7 --steps 2 --warmup 1 > $O/s22_bench_c3.json 2> $O/s22_bench_c3.err; el "bench c3 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/s22_bench_c3.json').read().strip().splitlines()[-1]); print('c3 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('max_rel_diff_vs_resident'))" | tee -a $O/s22_timeline.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py 3000 > $O/s22_sanitizer_memcheck.log 2>&1; el "memcheck rc=$?: $(tail -2 $O/s22_sanitizer_memcheck.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py 1500 > $O/s22_sanitizer_racecheck.log 2>&1; el "racecheck rc=$?: $(tail -2 $O/s22_sanitizer_racecheck.log | tr '\n' ' ')"
el done
