#!/bin/bash
# gpurun script: final state: full GPU suite, smoke(), default bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s28_timeline.log; }
timeout 900 python -m pytest tests -q -m gpu > $O/s28_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s28_pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/s28_smoke.log 2>&1; el "smoke rc=$?: $(tail -2 $O/s28_smoke.log | tr '\n' ' ')"
timeout 300 python bench.py > $O/s28_bench_n1.json 2> $O/s28_bench_n1.err; el "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/s28_bench_n1.json').read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'clocks', d['clocks'], 'cpu', d['cpu_baseline']['value'], 'double', d['double_precision']['kernel_ms'], d['double_precision']['parity_check'].get('exact'))" | tee -a $O/s28_timeline.log
el done
