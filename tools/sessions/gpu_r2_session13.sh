#!/bin/bash
# gpurun script: classified staging in the double-precision float-speed kernel (count_kernel_df): full GPU suite, bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s13_timeline.log; }
timeout 900 python -m pytest tests -q -m gpu -x > $O/s13_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s13_pytest_gpu.log)"
timeout 300 python bench.py > $O/s13_bench_n1.json 2> $O/s13_bench_n1.err; el "bench rc=$?"
python - <<'PY' | tee -a gpurun_out/s13_timeline.log
import json
d = json.loads(open('gpurun_out/s13_bench_n1.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d['parity_check'].get('within_reference_spread'), 'double', d.get('double_precision'))
PY
el done
