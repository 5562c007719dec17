#!/bin/bash
# gpurun script: streamed ingest (tests + timing against the one-shot upload), live peak measurements, bench line with roofline.ceiling.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s33_timeline.log; }
timeout 120 python -m pytest tests/test_gpu_stream_ingest.py tests/test_gpu_peaks.py -q -m gpu > $O/s33_pytest.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s33_pytest.log)"
timeout 60 python tools/time_ingest.py > $O/s33_ingest.jsonl 2> $O/s33_ingest.err; el "ingest rc=$?: $(cat $O/s33_ingest.jsonl | tr '\n' ' ')"
timeout 120 python bench.py > $O/s33_bench_n1.json 2> $O/s33_bench_n1.err; el "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/s33_bench_n1.json').read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'ceiling', d['roofline'].get('ceiling'))" | tee -a $O/s33_timeline.log
el done
