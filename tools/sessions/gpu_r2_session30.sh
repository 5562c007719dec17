#!/bin/bash
# gpurun script: configs[2] at its real size against the reference's sums (tests/golden/fullsize_c3.npz): GPU test + bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s30_timeline.log; }
timeout 600 python -m pytest tests/test_fullsize_golden.py -q -m gpu -k survey > $O/s30_pytest.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s30_pytest.log)"
timeout 600 python bench.py --workload c3_svy_spi_wt_2e6_2e7 --steps 2 --warmup 1 > $O/s30_bench_c3.json 2> $O/s30_bench_c3.err; el "bench c3 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/s30_bench_c3.json').read().strip().splitlines()[-1]); print('c3 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_check'])" | tee -a $O/s30_timeline.log
el done
