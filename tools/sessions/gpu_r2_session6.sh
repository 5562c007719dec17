#!/bin/bash
# gpurun script: device coordinate conversion tests, full GPU suite, default bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s6_timeline.log; }
timeout 600 python -m pytest tests/test_gpu_cnvt.py -q -m gpu -s > $O/s6_pytest_cnvt.log 2>&1; el "pytest cnvt rc=$?: $(tail -1 $O/s6_pytest_cnvt.log)"; grep ulp $O/s6_pytest_cnvt.log | tee -a $O/s6_timeline.log
timeout 900 python -m pytest tests -q -m gpu > $O/s6_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s6_pytest_gpu.log)"
timeout 300 python bench.py > $O/s6_bench_n1.json 2> $O/s6_bench_n1.err; el "bench rc=$?"
el done
