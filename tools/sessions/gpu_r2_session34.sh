#!/bin/bash
# gpurun script: state at the end of the round: full GPU suite, smoke(), default bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s34_timeline.log; }
timeout 900 python -m pytest tests -q -m gpu > $O/s34_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s34_pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/s34_smoke.log 2>&1; el "smoke rc=$?: $(tail -2 $O/s34_smoke.log | tr '\n' ' ')"
timeout 60 python tools/time_ingest.py > $O/s34_ingest.jsonl 2> $O/s34_ingest.err; el "ingest rc=$?: $(cat $O/s34_ingest.jsonl | tr '\n' ' ')"
timeout 300 python bench.py > $O/s34_bench_n1.json 2> $O/s34_bench_n1.err; el "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/s34_bench_n1.json').read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'clocks', d['clocks'], 'double', d['double_precision']['kernel_ms'], d['double_precision']['parity_check'].get('exact'))" | tee -a $O/s34_timeline.log
el done
