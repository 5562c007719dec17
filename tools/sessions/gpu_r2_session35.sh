#!/bin/bash
# gpurun script: compute-sanitizer memcheck over the streamed-ingest tests (pinned staging slots, growth of the device columns, errors).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s35_timeline.log; }
timeout 60 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stream_ingest.py -q -m gpu -k "edge or ascii or survey" > $O/s35_memcheck_stream.log 2>&1; el "memcheck rc=$?: $(grep -E 'ERROR SUMMARY|passed|failed' $O/s35_memcheck_stream.log | tr '\n' ' ')"
el done
